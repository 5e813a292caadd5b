"""GPU: oracle check of random unit samples drawn from the FULL-SIZE batches of the BASELINE.json configs (C2, C3, C5; C4 is in
test_gpu_tree_pipeline.py), evaluated exactly as bench.py evaluates them (bench.ConfigRunner: same models, same generator and
seed, same scenario chunks, same Jacobian-buffer reuse): for two chunks of every config the whole chunk runs through the
Jacobian pipeline, >= 5,000 randomly chosen units of it are re-computed by the CPU oracle (complex-step Jacobians) and compared
per state row and per Jacobian plane at 1e-9."""
import os
import sys

import numpy as np
import pytest

from conftest import ROOT, oracle_model_from_export, rel_err_rows

pytestmark = pytest.mark.gpu
TOL = 1e-9


@pytest.mark.parametrize("name,per_chunk", [("C2", 10000), ("C3", 5000), ("C5", 5000)])
def test_sample_of_the_full_size_batch_against_the_oracle(name, per_chunk):
    import torch
    sys.path.insert(0, ROOT)
    import bench
    from oracle.pyoracle import Oracle
    dev = torch.device("cuda", 0)
    r = bench.ConfigRunner(name, dev, 0, 1)
    cfg = bench.CONFIGS[name]
    assert r.B_local == cfg["B"] and r.N == cfg["N"] and r.B_local * r.N == r.U  # the config's own size, all of it on this GPU
    r.step()  # one full pass over every chunk, as the bench times it
    torch.cuda.synchronize()
    om = oracle_model_from_export(r.model)
    orc = Oracle(om, fast=True)
    n, P = r.n, 4 * r.n + 1
    rng = np.random.default_rng(2026)
    chunks = sorted({0, r.nchunks - 1})
    for ci in chunks:
        q, qd, tau, f = r.inp[ci]
        Uc = q.shape[1]
        # the Jacobian buffer holds the LAST chunk of the pass; re-run the sampled chunk into it (the states of every chunk
        # are still those of the pass and are checked as they are)
        states = [t.clone() for t in r.out[ci]]
        r.ev.step_rk4_jvp(q, qd, tau, f, cfg["dt"], out=r.out[ci], jac=r.jac)
        torch.cuda.synchronize()
        assert all(torch.equal(a, b) for a, b in zip(states, r.out[ci]))  # the pass and the re-run agree bit for bit
        idx = torch.from_numpy(np.sort(rng.choice(Uc, min(per_chunk, Uc), replace=False))).to(dev)
        hin = [np.ascontiguousarray(t[:, idx].cpu().numpy()) for t in (q, qd, tau, f)]
        got = [np.ascontiguousarray(t[:, idx].cpu().numpy()) for t in r.out[ci]]
        gj = r.jac[:, :, idx].cpu().numpy()
        if cfg.get("coupled"):
            f0, f1, w = r.model.coupling
            ref = orc.step_rk4_coupled((f0, f1), w, *hin, cfg["dt"], jac=True)
        else:
            ref = orc.step_rk4_jvp(*hin, cfg["dt"])
        for a, b, nm in zip(got, ref[:3], ("q+", "qd+", "f+")):
            assert rel_err_rows(a, b) < TOL, (name, ci, nm)
        err = rel_err_rows(gj.reshape(3 * n * P, -1), ref[3].reshape(3 * n * P, -1))
        assert err < TOL, (name, ci, err)
