#!/usr/bin/env python
"""Per-stage DRAM traffic, time and tensor-pipe activity of ONE exact-GP step from an ncu per-launch CSV.

The CSV comes from (GPU box, eager launches so that every kernel is its own ncu range):
  MOGP_GRAPH=0 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,\
sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed,sm__inst_executed_pipe_fp64.sum \
      --clock-control none --csv --log-file gpurun_out/step.csv python tools/one_step.py --config cfg2
Stages follow the launch order of capi.cu::enqueue_step: kbuild | potrf (+ pipelined inverse) | trtri | kinv | solves |
grad + finalize.  Per-launch times under ncu are serialised and cold-cache: compare SHARES, not absolutes.
Writes a JSON summary (bench.py reads profiles/r02_stage_traffic.json for `roofline.traffic`).
Usage: python tools/stage_traffic.py step.csv cfg2 [out.json]
"""
import csv
import json
import sys


def load(path):
    rows = []
    with open(path) as f:
        lines = [l for l in f if not l.startswith("==")]
    for r in csv.DictReader(lines):
        rows.append(r)
    per = {}
    order = []
    for r in rows:
        k = r["ID"]
        if k not in per:
            per[k] = {"name": r["Kernel Name"], "m": {}}
            order.append(k)
        try:
            v = float(r["Metric Value"].replace(",", ""))
        except ValueError:
            continue
        unit = r.get("Metric Unit", "")
        if r["Metric Name"] == "gpu__time_duration.sum":
            v *= {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(unit, 1e-3)       # -> microseconds
        if r["Metric Name"].startswith("dram__bytes"):
            v *= {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(unit, 1.0)
        per[k]["m"][r["Metric Name"]] = v
    return [per[k] for k in order]


def stage_of(name, state):
    n = name
    if "stage_inputs" in n or "prep_kernel" in n:
        return "kbuild"
    if "kbuild_kernel" in n:
        state["seen_kbuild"] = True
        return "kbuild"
    if "pad_copy" in n:                      # (staged before the factorisation since the early-loss change: no state change)
        return "solves"
    if "trmv_lower" in n or "colpass" in n or "lml_early" in n:
        state["solves"] = True
        return "solves"
    if "grad_reduce" in n or "pairsum" in n or "finalize" in n or "copy_out" in n or "params_" in n:
        return "grad_finalize"
    if state.get("solves"):
        return "kinv"
    if "diag_finish" in n:
        state["factored"] = True          # what follows (level-batched trtri_padded, int8 slicing + GEMMs) is the inverse
        return "potrf_inverse"
    return "trtri" if state.get("factored") else "potrf_inverse"


def main():
    path, cfg = sys.argv[1], sys.argv[2]
    out_path = sys.argv[3] if len(sys.argv) > 3 else None
    launches = load(path)
    # keep the LAST complete step: from the last prep/stage kernel before the last finalize
    last_fin = max(i for i, l in enumerate(launches) if "finalize_kernel" in l["name"])
    start = max(i for i, l in enumerate(launches[:last_fin]) if "prep_kernel" in l["name"])
    step = launches[start:last_fin + 1]
    state, stages = {}, {}
    # the K^-1 product is launched (on a side stream) before the solves: it is the last GEMM before trmv_lower -- together
    # with the three slicing kernels in front of it when it runs on the int8 pipe (i8_rowmax, i8_exponent, i8_slice_tiled)
    idx_pad = next((i for i, l in enumerate(step) if "trmv_lower" in l["name"]), None)
    kinv_set = set()
    if idx_pad is not None:
        i = idx_pad - 1
        while i >= 0 and "gemm" not in step[i]["name"]:
            i -= 1
        if i >= 0 and "i8_gemm" in step[i]["name"]:
            while i >= 0 and "i8_gemm" in step[i]["name"]:          # one launch (128 x 64 tiles) or two passes (128 x 128)
                kinv_set.add(i)
                i -= 1
            n = 0
            while i >= 0 and n < 3 and step[i]["name"].startswith("i8_") and "gemm" not in step[i]["name"]:
                kinv_set.add(i)                                       # i8_slice_tiled, i8_exponent, i8_rowmax of its operand
                i -= 1
                n += 1
        elif i >= 0:
            kinv_set.add(i)
    for i, l in enumerate(step):
        st = "kinv" if i in kinv_set else stage_of(l["name"], state)
        s = stages.setdefault(st, {"launches": 0, "time_us": 0.0, "dram_read": 0.0, "dram_write": 0.0, "tensor_weighted": 0.0,
                                   "kernels": {}})
        m = l["m"]
        t = m.get("gpu__time_duration.sum", 0.0)
        s["launches"] += 1
        s["time_us"] += t
        s["dram_read"] += m.get("dram__bytes_read.sum", 0.0)
        s["dram_write"] += m.get("dram__bytes_write.sum", 0.0)
        s["tensor_weighted"] += t * m.get("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", 0.0)
        kn = l["name"].split("(")[0][:60]
        kk = s["kernels"].setdefault(kn, [0, 0.0])
        kk[0] += 1
        kk[1] += t
    total = sum(s["time_us"] for s in stages.values())
    res = {"config": cfg, "source": path, "launches": len(step), "total_time_us_serialised": total, "stages": {}}
    for k, s in stages.items():
        res["stages"][k] = {"launches": s["launches"], "time_us": round(s["time_us"], 2), "share": round(s["time_us"] / total, 4),
                            "dram_bytes": s["dram_read"] + s["dram_write"], "dram_read": s["dram_read"], "dram_write": s["dram_write"],
                            "tensor_pipe_pct_time_weighted": round(s["tensor_weighted"] / max(s["time_us"], 1e-9), 2),
                            "kernels": {n: {"launches": c, "time_us": round(t, 2)} for n, (c, t) in s["kernels"].items()}}
    txt = json.dumps(res, indent=1)
    print(txt)
    if out_path:
        try:
            with open(out_path) as f:
                allr = json.load(f)
        except Exception:
            allr = {}
        allr[cfg] = res
        with open(out_path, "w") as f:
            json.dump(allr, f, indent=1)


if __name__ == "__main__":
    main()
