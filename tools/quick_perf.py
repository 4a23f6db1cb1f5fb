#!/usr/bin/env python3
"""Quick device-side timing of the MSM / NTT kernels (development aid; bench.py is the contract)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import aes_zero_knowledge_proof_circuit_b200 as zk

def main():
    logs = [int(x) for x in (sys.argv[1].split(",") if len(sys.argv) > 1 else ["16", "20", "22"])]
    windows = [int(x) for x in (sys.argv[2].split(",") if len(sys.argv) > 2 else ["0"])]
    # optional third argument: tuning sets separated by '/', each "key=value,key=value" (zkaes_ctx_set_tuning)
    tunings = sys.argv[3].split("/") if len(sys.argv) > 3 else [""]
    ctx = zk.Context(0)
    curve = 377
    stream = torch.cuda.ExternalStream(ctx.stream)
    for lg in logs:
        n = 1 << lg
        bases = torch.empty(n * 96, dtype=torch.uint8, device="cuda")
        t0 = time.time(); ctx.srs_powers_device(curve, bytes(range(32)), n, bases); ctx.sync()
        print(f"srs 2^{lg}: {time.time()-t0:.3f}s", flush=True)
        g = torch.Generator(device="cuda"); g.manual_seed(lg)
        sc = torch.randint(0, 2**62, (n, 4), dtype=torch.int64, device="cuda", generator=g)
        sc[:, 3] &= (1 << 59) - 1   # < r
        torch.cuda.synchronize()
        for tune in tunings:
          for kv in filter(None, tune.split(",")):
            ctx.set_tuning(kv.split("=")[0], int(kv.split("=")[1]))
          for w in windows:
            ctx.set_msm_window(w)
            for it in range(3):
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(stream); out = ctx.msm_g1_device(curve, bases, sc, n); e1.record(stream); e1.synchronize()
                ms = e0.elapsed_time(e1)
            print(f"msm 2^{lg} c={w} [{tune}]: {ms:.3f} ms  ({n/ms/1e3:.2f} Mpts/s, {128*n/ms/1e6:.1f} GB/s algorithmic)", flush=True)
        ctx.set_msm_window(0)
        data = sc.clone()
        for it in range(3):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream); ctx.ntt_fr_device(curve, data, lg); e1.record(stream); e1.synchronize()
            ms = e0.elapsed_time(e1)
        print(f"ntt 2^{lg}: {ms:.3f} ms ({64*n/ms/1e6:.1f} GB/s algorithmic)", flush=True)
        del bases, sc, data
main()
