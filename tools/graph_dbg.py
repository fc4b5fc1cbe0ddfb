import sys, os
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/tests')
import torch, numpy as np
from conftest import load_golden
from mogptk_b200.engine import Engine, pack_params
eng = Engine(0, 8192)
def run(name, n=4, want=True):
    g = load_golden(name)
    rows = eng.prepare(g["kind"], g["params"], g["X"], g["y"])
    p = pack_params(g["kind"], g["params"], eng.device)
    sig = torch.tensor(g["sigma"], device=eng.device)
    for i in range(n):
        out = eng.lml_grad_prepared(rows, p, sig, g["jitter"], want, check=False)
        torch.cuda.synchronize()
        print(name, i, "lml", float(out[0]), "info", float(out[1]), "ref", float(g["lml"]), "g0", float(out[2]) if want else None, flush=True)
run("cfg2"); run("mosm_mid"); run("cfg3", 2); run("cfg2"); run("cfg4", 3); run("cfg2", 3, False); run("cfg2", 3)
