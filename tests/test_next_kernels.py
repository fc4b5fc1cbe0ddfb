"""CPU tests for the next kernel families on the hot path (CSM, SM-LMC): the oracle restatement
reproduces the reference's K (golden fixtures written by oracle/make_golden_next.py from the live reference), and the
per channel-pair component table in the product's one derived form reproduces it too -- i.e. these families need a new
table (csrc/covmath.cuh) but no new CUDA kernels."""
import os

import numpy as np
import pytest
import torch

from conftest import GOLDEN_DIR, next_golden_names
from oracle import next_kernels as nk


def _load(name):
    z = np.load(os.path.join(GOLDEN_DIR, name + ".npz"), allow_pickle=False)
    g = {k: z[k] for k in z.files}
    p = {k[2:]: torch.tensor(v, dtype=torch.float64) for k, v in g.items() if k.startswith("p_")}
    X = torch.tensor(g["X"], dtype=torch.float64)
    C = int(g["C"])
    rows = [torch.nonzero(X[:, 0].long() == c, as_tuple=False)[:, 0] for c in range(C)]
    return str(g["kind"]), C, p, X, rows, torch.tensor(g["K"], dtype=torch.float64)


@pytest.mark.parametrize("name", next_golden_names())
def test_restatement_matches_the_reference(name):
    kind, C, p, X, rows, K = _load(name)
    for i in range(C):
        for j in range(C):
            blk = nk.KSUB[kind](i, j, X[rows[i], 1:], X[rows[j], 1:], p)
            assert float((blk - K[rows[i]][:, rows[j]]).abs().max()) <= 1e-13 * float(K.abs().max())


@pytest.mark.parametrize("name", next_golden_names())
def test_derived_component_form_matches_the_reference(name):
    kind, C, p, X, rows, K = _load(name)
    for i in range(C):
        for j in range(C):
            comps = nk.derived_components(kind, p, i, j)
            blk = nk.k_from_components(comps, X[rows[i], 1:], X[rows[j], 1:])
            assert float((blk - K[rows[i]][:, rows[j]]).abs().max()) <= 1e-13 * float(K.abs().max())
    # the table is symmetric in the sense the Gram build relies on: block (j, i) is the transpose of block (i, j)
    for i in range(C):
        for j in range(i):
            a = nk.k_from_components(nk.derived_components(kind, p, i, j), X[rows[i], 1:], X[rows[j], 1:])
            b = nk.k_from_components(nk.derived_components(kind, p, j, i), X[rows[j], 1:], X[rows[i], 1:])
            assert float((a - b.T).abs().max()) <= 1e-14 * float(K.abs().max())


def test_fixtures_exist():
    assert len(next_golden_names()) >= 4
