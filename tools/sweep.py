#!/usr/bin/env python3
"""BASELINE.json configs[4]: G1 MSM + Fr NTT sweep 2^16..2^24 on one B200 next to the CPU oracle (all host threads).

    python tools/sweep.py [curve=381] [max_log=24] [cpu_max_log=20]

GPU numbers: device-resident inputs, CUDA events on the library's stream, best of 3 after a warm-up.  The GPU result of every
MSM up to cpu_max_log is compared with the oracle's (bit-exact) -- the sweep doubles as a parity run at sizes the unit tests
do not reach.  CPU numbers: oracle/zk_oracle.cpp (restatement of ark-ec 0.3.0 Pippenger / ark-poly radix-2 FFT), one run.
Prints a markdown table (committed under profiles/)."""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

import aes_zero_knowledge_proof_circuit_b200 as zk
from oracle.cpu import Oracle, rand_fr


def main():
    curve = int(sys.argv[1]) if len(sys.argv) > 1 else 381
    max_log = int(sys.argv[2]) if len(sys.argv) > 2 else 24
    cpu_max = int(sys.argv[3]) if len(sys.argv) > 3 else 20
    orc = Oracle()
    ctx = zk.Context(0)
    stream = torch.cuda.ExternalStream(ctx.stream)
    print(f"# BLS12-{curve} G1 MSM / Fr NTT sweep, one B200 vs CPU oracle ({orc.threads()} host threads)")
    print("| log2 n | GPU MSM ms | Mterms/s | CPU MSM ms | MSM speed-up | parity | GPU NTT ms | CPU NTT ms | NTT speed-up |")
    print("|---|---|---|---|---|---|---|---|---|")

    def gpu_time(fn):
        best = None
        for it in range(4):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            out = fn()
            e1.record(stream)
            e1.synchronize()
            if it:
                best = e0.elapsed_time(e1) if best is None else min(best, e0.elapsed_time(e1))
        return best, out

    for lg in range(16, max_log + 1, 2):
        n = 1 << lg
        bases = torch.empty(n * 96, dtype=torch.uint8, device="cuda")
        ctx.srs_powers_device(curve, bytes(range(32)), n, bases)
        ctx.sync()
        sc_host = rand_fr(np.random.default_rng(lg), curve, n)
        sc = torch.from_numpy(sc_host.view(np.int64)).cuda()
        msm_ms, got = gpu_time(lambda: ctx.msm_g1_device(curve, bases, sc, n))
        data = torch.from_numpy(sc_host.view(np.int64)).cuda()
        ntt_ms, _ = gpu_time(lambda: ctx.ntt_fr_device(curve, data, lg))
        cpu_msm = cpu_ntt = parity = "-"
        if lg <= cpu_max:
            bh = bases.cpu().numpy().view(np.uint64).reshape(n, 12)
            t0 = time.perf_counter()
            exp = orc.g1_msm(curve, bh, sc_host)
            cpu_msm = (time.perf_counter() - t0) * 1e3
            parity = "bit-exact" if (got == exp).all() else "MISMATCH"
            t0 = time.perf_counter()
            orc.ntt(curve, sc_host)
            cpu_ntt = (time.perf_counter() - t0) * 1e3
        f = lambda v: f"{v:.2f}" if isinstance(v, float) else v
        sp = lambda c, g: f"{c / g:.0f}x" if isinstance(c, float) else "-"
        print(f"| {lg} | {msm_ms:.2f} | {n / msm_ms / 1e3:.1f} | {f(cpu_msm)} | {sp(cpu_msm, msm_ms)} | {parity} | {ntt_ms:.3f} | {f(cpu_ntt)} | {sp(cpu_ntt, ntt_ms)} |",
              flush=True)
        del bases, sc, data
    ctx.close()


main()
