#!/usr/bin/env python3
"""bench.py -- headline benchmark of the B200-native prover path.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--msg-len BYTES] [--workload prove|msm|ntt] [--curve 377|381] [--log-n L]
                    [--impl reference]

Metric (BASELINE.json): encrypt() prove time and constraints/s.  One "step" = one encrypt(): AES-128-ECB witness
generation + Marlin proof of one synthetic message through the C ABI (host buffers in, ciphertext + proof bytes out).
The proving key (test SRS, matrices, index polynomials) is resident in HBM before the timed region, exactly as the
reference passes an already-synthesised ProvingKey to encrypt() (src/lib.rs:60-64).

    value        constraints/s, timed with CUDA events on the library's stream (host orchestration between kernels included)
    e2e          the same metric by wall clock around the C-ABI call: message + key H2D and ciphertext + proof D2H inside
    roofline     the dominant kernel, the MSM bucket accumulation (k_msm_accumulate): bracketed by CUDA events inside the
                 library during the timed steps; achieved = 128 B x MSM terms / kernel time
    config.verified / proof_sha256   the LAST timed proof is checked with verify_encryption (host pairing verifier) and hashed,
                 so the 1/2/4/8-GPU lines show that they computed the same, valid proof
    cpu_baseline ONE configuration both arms share: the full 16-byte proof of BASELINE.json configs[0] (src/main.rs:9-26;
                 185,040 constraints, |H| = 2^18, |K| = 2^19) by the CPU oracle prover (oracle/marlin_oracle.py: ark-marlin
                 0.3.0 restated over the oracle's C++ MSM/NTT, all host threads), with `gpu_same_config_ms` = the same
                 16-byte proof on this GPU next to it.  The 4 KiB headline is NOT run on the CPU (its SRS alone is 39 GB);
                 `extrapolation` states how the 16-byte pair relates to it.
--impl reference times that CPU prover, one full 16-byte proof per step (the reference itself is Rust with un-vendored
crates; this image has no cargo, so the oracle port is the only CPU implementation of the path here).  A CPU proof takes
tens of seconds, so the arm stops after --cpu-budget-s seconds of timed steps (default 240) and reports the steps it timed.

--workload msm / ntt (BASELINE.json configs[4]): one 2^log_n-term G1 MSM / one 2^log_n-point Fr NTT per step on BLS12-377 or
BLS12-381; MSM shards by point range over the ranks (NTT is per GPU: every rank runs the same transform); --cpu-check adds the
CPU oracle's time for the same size and, for the MSM, a bit-exact comparison of the results.
"""
import argparse
import hashlib
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
# stdout carries exactly one JSON line.  NCCL prints its banner / debug output to fd 1 from native code (seen on the GPU box:
# "NCCL version ..." ahead of the JSON), so fd 1 is pointed at stderr for the whole process and the result line is written to
# a private duplicate of the original stdout.
sys.stdout.flush()
_RESULT_FD = os.dup(1)
os.dup2(2, 1)


def emit(line: dict):
    os.write(_RESULT_FD, (json.dumps(line) + "\n").encode())

SEED_TAU, SEED_GAMMA, SEED_ZK = bytes(range(32)), bytes(range(1, 33)), bytes([7] * 32)
AES_KEY = bytes.fromhex("2b7e151628aed2a6abf7158809cf4f3c")  # FIPS-197 (SURVEY.md 8(d) synthetic inputs)
# Integer-pipe peak of one B200, tools/ubench.cu (profiles/ubench_r2.txt): 9.05e12 IMAD.WIDE/s (31.1 /clk/SM x 148 SMs x 1965 MHz; every
# 32x32->64 multiply-add runs at half the IMAD rate on sm_100).  A mixed addition is 8 Fq products x 288 IMAD.WIDE (12 x 12 partial
# products + 12 x 12 reduction) + 2 dedicated squarings x 222 (78 + 144) = 2748, so the pipe allows 9.05e12 / 2748 = 3.29 G mixed
# additions/s; that is the denominator of alu.frac.
IMAD_WIDE_PEAK_PER_S = 9.05e12
IMAD_WIDE_PER_MADD = 8 * 288 + 2 * 222
MADD_PEAK_PER_S = IMAD_WIDE_PEAK_PER_S / IMAD_WIDE_PER_MADD
MADD_UBENCH_PER_S = 2.46e9  # the same addition formula in a register-resident microbenchmark loop (tools/ubench.cu k_madd)


def ncu_traffic_per_term():
    """DRAM bytes per MSM term of one k_msm_accumulate launch (dram__bytes_read.sum + dram__bytes_write.sum of an `ncu --set full`
    capture / terms of that launch), read from the committed summary of this round's capture -- not a constant in this file."""
    path = os.path.join(ROOT, "profiles", "r2_ncu_traffic.json")
    try:
        with open(path) as f:
            d = json.load(f)
        return float(d["dram_bytes_per_term"]), d.get("source", path)
    except (OSError, KeyError, ValueError):
        return None, None


def synth_message(n):
    return bytes((i * 131 + 7) & 0xFF for i in range(n))


def load_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            d = json.load(f)
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json, burst copy)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clock / throttle-reason sampler running during the timed region."""

    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index = index
        self.rows = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        for r in self.rows:
            try:
                sm.append(float(r[0]))
                mx.append(float(r[1]))
            except (ValueError, IndexError):
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ------------------------------------------------------------------------------------------------------------
# CPU side: the oracle prover on the one configuration both arms can run -- the full 16-byte circuit (BASELINE configs[0])
# ------------------------------------------------------------------------------------------------------------
SHARED_MSG = bytes.fromhex("3243f6a8885a308d313198a2e0370734")  # the reference's own end-to-end case (tests/integration_tests.rs:313-337)
SHARED_DESC = "encrypt() prove, 16-byte message (1 ECB block): BASELINE.json configs[0] (src/main.rs:9-26), AES-128-ECB R1CS, Marlin/BLS12-377"
EXTRAPOLATION = ("16 B -> 4 KiB: constraints x205 (185,040 -> 37,994,400) while the domains the prover's MSMs / NTTs run over grow x256 "
                 "(|H| 2^18 -> 2^26, |K| 2^19 -> 2^27); a CPU Pippenger at the larger size uses c ~ 21 instead of ~ 17 (windows 15 -> 12), so CPU "
                 "time at 4 KiB ~ 256 x 0.8 x the 16-byte time.  An estimate: the 4 KiB proof itself was never run on the CPU (SRS 39 GB, ~1.5 h).")


class CpuSample:
    """The full 16-byte AES R1CS (185,040 constraints), indexed once by the oracle; each prove() is one complete Marlin proof."""

    def __init__(self, prefix=0):
        from oracle import marlin_oracle as mo
        from oracle import r1cs_model as model

        self.mo = mo
        t0 = time.perf_counter()
        cs, self.ct = model.synthesize(SHARED_MSG, AES_KEY)
        A, B, C = cs.matrices()
        ninst, nwit = len(cs.inst_vals), len(cs.wit_vals)
        self.inst, self.wit = cs.inst_vals, cs.wit_vals
        self.prefix = prefix
        if prefix:
            # --cpu-sample-constraints (the CPU-tier tests' quick mode): only the first `prefix` constraints.  Every constraint touches
            # earlier variables only, so the prefix with the real wire values is a satisfied R1CS of its own (witnesses renumbered).
            A, B, C = A[:prefix], B[:prefix], C[:prefix]
            top = max(c for m in (A, B, C) for row in m for c, _ in row)
            assert all(c == 0 or c >= ninst for m in (A, B, C) for row in m for c, _ in row)
            nwit, shift = top - ninst + 1, ninst - 1
            fix = lambda m: [[(c - shift if c >= ninst else c, v) for c, v in row] for row in m]
            A, B, C, ninst = fix(A), fix(B), fix(C), 1
            self.inst, self.wit = [1], list(cs.wit_vals[:nwit])
        self.r1cs = mo.R1CS(A, B, C, ninst, nwit)
        self.n_constraints = len(A)
        idx0 = mo.index_r1cs(self.r1cs)
        self.srs = mo.SRS.generate(idx0.max_degree, SEED_TAU, SEED_GAMMA)
        self.idx = mo.index_r1cs(self.r1cs, self.srs)
        self.cores = mo.orc().threads()
        self.setup_s = time.perf_counter() - t0

    def prove(self):
        t0 = time.perf_counter()
        _, pb = self.mo.prove(self.idx, self.srs, self.r1cs, self.inst, self.wit, SEED_ZK)
        return time.perf_counter() - t0, pb

    def describe(self):
        what = f"the first {self.prefix} constraints of the 16-byte AES-128 circuit" if self.prefix else "the 16-byte AES-128 circuit"
        return (f"one full Marlin proof of {what} ({self.n_constraints} constraints, |H|={self.idx.domain_h.size}, "
                f"|K|={self.idx.domain_k.size}); CPU restatement of ark-marlin 0.3.0 (oracle port), std::thread MSM/NTT on all cores")


def run_reference(args, rank):
    if rank != 0:
        return
    s = CpuSample(args.cpu_sample_constraints)
    times, proof = [], b""
    for _ in range(min(args.warmup, 1)):  # a CPU proof has no clocks / caches to warm beyond the first call
        s.prove()
    t_start = time.perf_counter()
    for _ in range(args.steps):
        dt, proof = s.prove()
        times.append(dt)
        if time.perf_counter() - t_start > args.cpu_budget_s and len(times) >= 2:
            break
    dt = float(np.mean(times))
    val = s.n_constraints / dt
    line = {
        "impl": "reference", "metric": "encrypt_prove_constraints_per_s", "value": val, "unit": "constraints/s", "n_gpus": args.gpus,
        "steps": len(times), "steps_requested": args.steps, "warmup": min(args.warmup, 1), "ms_per_step": dt * 1e3, "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "u64x4 / u64x6 Montgomery (BLS12-377 Fr / Fq integers)", "data": "synthetic",
        "config": {"workload": SHARED_DESC if not s.prefix else f"PREFIX SAMPLE (first {s.prefix} constraints) of: " + SHARED_DESC, "msg_len": 16, "constraints": s.n_constraints, "H": s.idx.domain_h.size, "K": s.idx.domain_k.size,
                   "proof_sha256": hashlib.sha256(proof).hexdigest(), "key_setup_s": s.setup_s,
                   "note": "the GPU arm's headline config is the 4 KiB message; this arm runs the largest configuration a CPU finishes in "
                           "seconds, and the GPU arm reports the same 16-byte proof as cpu_baseline.gpu_same_config_ms",
                   "extrapolation": EXTRAPOLATION},
        "cpu_baseline": {"value": val, "unit": "constraints/s", "cores": s.cores, "kind": "port", "sample": s.describe()},
        "e2e": {"value": val, "unit": "constraints/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    emit(line)


# ------------------------------------------------------------------------------------------------------------
def bench_prove(args, rank, world, local_rank):
    import torch
    import torch.distributed as dist

    import aes_zero_knowledge_proof_circuit_b200 as zk

    ctx = zk.Context(local_rank)
    if world > 1:
        # the prover's MSMs shard by point range over the ranks; the library runs its own NCCL all-gather of window sums
        uid = [zk.comm_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(uid, src=0)
        ctx.comm_init(rank, world, uid[0])
    stream = torch.cuda.ExternalStream(ctx.stream)
    msg_len = args.msg_len
    msg = synth_message(msg_len)
    t0 = time.perf_counter()
    pk = ctx.synthesize_keys(msg_len, SEED_TAU, SEED_GAMMA)
    setup_s = time.perf_counter() - t0
    n_constraints = pk.info["num_constraints"]

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        ctx.sync()

    for _ in range(args.warmup):
        ct, proof = ctx.encrypt(pk, msg, AES_KEY, SEED_ZK)
    barrier()
    l0 = ctx.launches
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    ctx.profile(True)
    ctx.profile_read()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record(stream)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        ct, proof = ctx.encrypt(pk, msg, AES_KEY, SEED_ZK)
    e1.record(stream)
    e1.synchronize()
    barrier()
    wall_ms = (time.perf_counter() - t0) * 1e3
    dev_ms = e0.elapsed_time(e1)
    prof = ctx.profile_read()
    ctx.profile(False)
    clocks = sampler.stop() if rank == 0 else None
    launches = ctx.launches - l0
    if world > 1:
        t = torch.tensor([dev_ms, wall_ms], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dev_ms, wall_ms = float(t[0].item()), float(t[1].item())
    if rank != 0:
        pk.close()
        return None
    ms_per_step = dev_ms / args.steps
    e2e_ms = wall_ms / args.steps
    # one proof per step at every N: the MSMs are sharded over the ranks (strong scaling); witness, NTTs and the
    # transcript are computed redundantly by every rank
    units = n_constraints
    peak, peak_src = load_peaks()
    acc_ms = prof["ms"] / max(prof["launches"], 1)
    alg_bytes = 128.0 * prof["terms"] / max(prof["launches"], 1)
    achieved = alg_bytes / (acc_ms * 1e-3) / 1e9 if acc_ms > 0 else 0.0
    madds_per_s = prof["madds"] / (prof["ms"] * 1e-3) if prof["ms"] else 0.0
    bpt, bpt_src = ncu_traffic_per_term()
    verified = bool(zk.verify_encryption(pk.verifying_key(), proof, ct))  # the last timed proof, host pairing verifier
    line = {
        "metric": "encrypt_prove_constraints_per_s", "value": units / (ms_per_step * 1e-3), "unit": "constraints/s", "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None,
        "dtype": "u32x8 / u32x12 Montgomery (BLS12-377 Fr / Fq integers)", "data": "synthetic",
        "config": {"workload": f"encrypt() prove, {msg_len}-byte message ({msg_len // 16} ECB blocks), AES-128-ECB R1CS, Marlin/BLS12-377",
                   "msg_len": msg_len, "constraints": n_constraints, "H": pk.info["h"], "K": pk.info["k"], "srs_points": pk.info["max_degree"] + 1,
                   "parallelism": "single GPU" if world == 1 else (f"MSM point-range x{world} (NCCL all-gather of window sums), coset transforms spread over the ranks, "
                                                                     "w / z_A / z_B chains one owner each, r_alpha and f in slices; witness and transcript replicated"),
                   "r1_lagrange_basis": bool(pk.info.get("lagrange_points")),
                   "cache": "per-step working set (index polynomials + SRS + round buffers) exceeds the 126 MB L2; no flush needed",
                   "key_setup_s": setup_s, "proof_bytes": len(proof), "proof_sha256": hashlib.sha256(proof).hexdigest(),
                   "ciphertext_sha256": hashlib.sha256(ct).hexdigest(), "verified": verified},
        "roofline": {"bound": "hbm", "kernel": "k_msm_accumulate (MSM bucket accumulation, XYZZ += affine)",
                     "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                     "traffic": (bpt * prof["terms"] / max(prof["launches"], 1)) if bpt else None,
                     "traffic_note": (f"bytes per launch = ncu-measured DRAM bytes per MSM term ({bpt_src}) x terms per launch; the bucket method gathers "
                                      "every point once per window (DESIGN.md 4.6)") if bpt else "no ncu capture of this round's kernel committed",
                     "peak_source": peak_src, "launches_per_step": prof["launches"] / args.steps, "avg_launch_ms": acc_ms, "algorithmic_bytes_per_launch": alg_bytes,
                     "share_of_step": prof["ms"] / dev_ms if dev_ms else None,
                     "alu": {"madds_per_s": madds_per_s, "imad_wide_per_s": madds_per_s * IMAD_WIDE_PER_MADD, "imad_wide_peak_per_s": IMAD_WIDE_PEAK_PER_S,
                             "frac": madds_per_s / MADD_PEAK_PER_S, "madd_peak_per_s": MADD_PEAK_PER_S,
                             "frac_of_microbenchmark": madds_per_s / MADD_UBENCH_PER_S,
                             "note": "the kernel is integer-multiplier bound (8 Fq products + 2 squarings = 2748 IMAD.WIDE per mixed addition): frac = IMAD.WIDE issued per "
                                     "second / the measured IMAD.WIDE pipe peak; the HBM fraction above is reported as the contract asks"}},
        "e2e": {"value": units / (e2e_ms * 1e-3), "unit": "constraints/s", "ms_per_step": e2e_ms, "h2d_bytes_per_step": msg_len + 16 + 32,
                "d2h_bytes_per_step": msg_len + len(proof)},
        "gpu_launches": int(launches), "clocks": clocks,
    }
    pk.close()
    if world == 1 and not args.no_cpu_baseline:
        # the configuration the CPU arm runs (full 16-byte proof), on this GPU: untimed by the headline, reported next to the CPU number
        pk16 = ctx.synthesize_keys(16, SEED_TAU, SEED_GAMMA)
        for _ in range(3):
            ct16, proof16 = ctx.encrypt(pk16, SHARED_MSG, AES_KEY, SEED_ZK)
        ctx.sync()
        t0 = time.perf_counter()
        reps = 10
        for _ in range(reps):
            ct16, proof16 = ctx.encrypt(pk16, SHARED_MSG, AES_KEY, SEED_ZK)
        ctx.sync()
        line["_gpu16"] = {"ms": (time.perf_counter() - t0) * 1e3 / reps, "proof_sha256": hashlib.sha256(proof16).hexdigest(),
                          "constraints": pk16.info["num_constraints"],
                          "verified": bool(zk.verify_encryption(pk16.verifying_key(), proof16, ct16))}
        pk16.close()
    return line


FR_TOP_LIMB = {377: 0x12ab655e9a2ca556, 381: 0x73eda753299d7d48}  # top 64 bits of the scalar moduli


def rand_scalars(curve, n, seed):
    rng = np.random.default_rng(seed)
    a = rng.integers(0, 1 << 63, size=(n, 4), dtype=np.uint64) * np.uint64(2) + rng.integers(0, 2, size=(n, 4), dtype=np.uint64)
    a[:, 3] %= np.uint64(FR_TOP_LIMB[curve])
    return a


def bench_msm(args, rank, world, local_rank):
    """BASELINE.json configs[4]: one 2^log_n-term G1 MSM per step; ranks hold contiguous point ranges, window sums are all-gathered (NCCL)
    and folded.  Inputs resident in HBM; result (one affine point) returned to the host every step."""
    import torch
    import torch.distributed as dist

    import aes_zero_knowledge_proof_circuit_b200 as zk
    curve = args.curve
    ctx = zk.Context(local_rank)
    ctx.set_tuning("msm_plan_ranks", world)  # the window plan of the stand-alone entry points should fit the per-rank share
    log_n = args.log_n
    n_total = 1 << log_n
    n_local = n_total // world
    lo = rank * n_local
    all_bases = torch.empty(n_total * 96, dtype=torch.uint8, device="cuda")
    ctx.srs_powers_device(curve, SEED_TAU, n_total, all_bases)
    ctx.sync()
    bases = all_bases[lo * 96:(lo + n_local) * 96].clone()
    scal_all = rand_scalars(curve, n_total, 2024)
    scalars = torch.from_numpy(np.ascontiguousarray(scal_all[lo:lo + n_local]).view(np.int64)).cuda()
    cpu = None
    if args.cpu_check and rank == 0:
        from oracle.cpu import Oracle
        orc = Oracle()
        bh = all_bases.cpu().numpy().view(np.uint64).reshape(n_total, 12)
        t0 = time.perf_counter()
        exp = orc.g1_msm(curve, bh, scal_all)
        cpu = {"ms": (time.perf_counter() - t0) * 1e3, "cores": orc.threads(), "expected": exp,
               "what": "oracle/zk_oracle.cpp: ark-ec 0.3.0 Pippenger restated, std::thread over windows"}
    del all_bases
    wbytes = ctx.msm_g1_windows_bytes(curve, n_total)
    win = torch.zeros(wbytes, dtype=torch.uint8, device="cuda")
    gathered = torch.zeros(world * wbytes, dtype=torch.uint8, device="cuda")
    torch.cuda.synchronize()

    def step():
        ctx.msm_g1_windows(curve, bases, scalars, n_local, n_total, win)
        if world > 1:
            ctx.sync()
            dist.all_gather_into_tensor(gathered, win)
            torch.cuda.synchronize()
            return ctx.msm_g1_fold(curve, gathered, world, n_total)
        return ctx.msm_g1_fold(curve, win, 1, n_total)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        ctx.sync()

    for _ in range(args.warmup):
        out = step()
    barrier()
    l0 = ctx.launches
    ctx.profile(True)
    ctx.profile_read()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        out = step()
    barrier()
    ms = (time.perf_counter() - t0) * 1e3
    prof = ctx.profile_read()
    clocks = sampler.stop() if rank == 0 else None
    launches = ctx.launches - l0
    if world > 1:
        t = torch.tensor([ms], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    if rank != 0:
        return None
    ms_per_step = ms / args.steps
    peak, peak_src = load_peaks()
    acc_ms = prof["ms"] / max(prof["launches"], 1)
    alg_bytes = 128.0 * prof["terms"] / max(prof["launches"], 1)
    achieved = alg_bytes / (acc_ms * 1e-3) / 1e9 if acc_ms > 0 else 0.0
    madds_per_s = prof["madds"] / (prof["ms"] * 1e-3) if prof["ms"] else 0.0
    plan_w = wbytes // 192
    line = {
        "metric": "msm_terms_per_s", "value": n_total / (ms_per_step * 1e-3), "unit": "G1 terms/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": f"u32x12 Montgomery (Fq, {curve}-bit integer)", "data": "synthetic",
        "config": {"workload": f"BLS12-{curve} G1 MSM 2^{log_n}", "curve": curve, "log_n": log_n, "windows": plan_w,
                   "parallelism": f"point-range x{world}, NCCL all-gather of window sums",
                   "cache": f"{128 * n_local >> 20} MiB of bases+scalars per pass" + (" > 126 MB L2" if 128 * n_local > 126 << 20 else " (fits L2)"),
                   "result_sha256": hashlib.sha256(np.asarray(out).tobytes()).hexdigest()},
        "roofline": {"bound": "hbm", "kernel": "k_msm_accumulate", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                     "traffic": None, "peak_source": peak_src, "avg_launch_ms": acc_ms, "share_of_step": prof["ms"] / ms if ms else None,
                     "alu": {"madds_per_s": madds_per_s, "madd_peak_per_s": MADD_PEAK_PER_S, "frac": madds_per_s / MADD_PEAK_PER_S,
                             # whole-MSM view: the mixed additions the window plan needs (n x W) at the pipe's peak / the measured step
                             "frac_whole_msm": (n_total / world * plan_w / MADD_PEAK_PER_S) / (ms_per_step * 1e-3)}},
        "e2e": {"value": None, "unit": "G1 terms/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 96, "note": "sweep workload: inputs device-resident"},
        "gpu_launches": int(launches), "clocks": clocks,
    }
    if cpu:
        line["cpu_baseline"] = {"value": n_total / (cpu["ms"] * 1e-3), "unit": "G1 terms/s", "cores": cpu["cores"], "kind": "port", "sample": cpu["what"],
                                "cpu_ms": cpu["ms"], "bit_exact": bool((np.asarray(out) == cpu["expected"]).all())}
    return line


def bench_ntt(args, rank, world, local_rank):
    """BASELINE.json configs[4]: one forward 2^log_n-point Fr NTT per step, in place in HBM.  The path does not shard (north_star: "NTT ...
    stays per-GPU"): with N ranks every rank runs the same transform (replicas), value = N transforms' points / time."""
    import torch
    import torch.distributed as dist

    import aes_zero_knowledge_proof_circuit_b200 as zk
    curve = args.curve
    ctx = zk.Context(local_rank)
    stream = torch.cuda.ExternalStream(ctx.stream)
    log_n = args.log_n
    n = 1 << log_n
    host = rand_scalars(curve, n, 7 + log_n)
    data = torch.from_numpy(host.view(np.int64)).cuda()
    cpu = None
    if args.cpu_check and rank == 0:
        from oracle.cpu import Oracle
        orc = Oracle()
        t0 = time.perf_counter()
        exp = orc.ntt(curve, host)
        cpu = {"ms": (time.perf_counter() - t0) * 1e3, "cores": orc.threads(), "expected": exp}
        ctx.ntt_fr_device(curve, data, log_n)
        ctx.sync()
        cpu["bit_exact"] = bool((data.cpu().numpy().view(np.uint64).reshape(n, 4) == exp).all())

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        ctx.sync()

    for _ in range(args.warmup):
        ctx.ntt_fr_device(curve, data, log_n)
    barrier()
    l0 = ctx.launches
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record(stream)
    for _ in range(args.steps):
        ctx.ntt_fr_device(curve, data, log_n)
    e1.record(stream)
    e1.synchronize()
    barrier()
    ms = e0.elapsed_time(e1)
    clocks = sampler.stop() if rank == 0 else None
    launches = ctx.launches - l0
    if world > 1:
        t = torch.tensor([ms], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    if rank != 0:
        return None
    ms_per_step = ms / args.steps
    peak, peak_src = load_peaks()
    achieved = 64.0 * n / (ms_per_step * 1e-3) / 1e9
    # n/2 log2 n butterflies x one Fr product of 128 IMAD.WIDE (8 x 8 + 8 x 8): the integer floor of the transform
    int_floor_ms = (n / 2 * log_n * 128) / IMAD_WIDE_PEAK_PER_S * 1e3
    line = {
        "metric": "ntt_points_per_s", "value": world * n / (ms_per_step * 1e-3), "unit": "Fr points/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "u32x8 Montgomery (Fr integer)", "data": "synthetic",
        "config": {"workload": f"BLS12-{curve} Fr NTT 2^{log_n} (forward, in place)", "curve": curve, "log_n": log_n, "parallelism": f"{world} replicas (no collective)",
                   "cache": f"{32 * n >> 20} MiB in place" + (" > 126 MB L2" if 32 * n > 126 << 20 else " (fits L2: numbers below 2^22 are L2-resident)")},
        "roofline": {"bound": "hbm", "kernel": "k_ntt_pass (all passes of one transform)", "achieved": achieved, "peak": peak, "unit": "GB/s",
                     "frac": achieved / peak, "traffic": None, "peak_source": peak_src,
                     "alu": {"imad_wide_floor_ms": int_floor_ms, "frac": int_floor_ms / ms_per_step}},
        "e2e": {"value": None, "unit": "Fr points/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0, "note": "sweep workload: data device-resident"},
        "gpu_launches": int(launches), "clocks": clocks,
    }
    if cpu:
        line["cpu_baseline"] = {"value": n / (cpu["ms"] * 1e-3), "unit": "Fr points/s", "cores": cpu["cores"], "kind": "port",
                                "sample": "oracle/zk_oracle.cpp radix-2 NTT (ark-poly 0.3.0 restated), std::thread", "cpu_ms": cpu["ms"], "bit_exact": cpu["bit_exact"]}
    return line


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--workload", default="prove", choices=["prove", "msm", "ntt"])
    ap.add_argument("--msg-len", type=int, default=4096)  # BASELINE.json: the metric is quoted on the 4 KiB message
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--log-n", type=int, default=22)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--cpu-budget-s", type=float, default=240.0)  # --impl reference: stop timing further CPU proofs after this many seconds
    ap.add_argument("--cpu-sample-constraints", type=int, default=0)  # --impl reference quick mode for the CPU-tier tests: a prefix of the circuit
    ap.add_argument("--curve", type=int, default=377, choices=[377, 381])  # msm / ntt workloads (the prover is BLS12-377, src/lib.rs:47)
    ap.add_argument("--cpu-check", action="store_true")  # msm / ntt: time the CPU oracle on the same size; MSM results compared bit for bit
    args = ap.parse_args()
    if args.impl == "b200":
        args.warmup = max(args.warmup, 3)  # timing rule: at least three untimed steps before a device measurement

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank)
        return

    import torch
    import torch.distributed as dist

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- the prover path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    line = {"prove": bench_prove, "msm": bench_msm, "ntt": bench_ntt}[args.workload](args, rank, world, local_rank)
    if rank == 0:
        if args.workload == "prove" and world == 1 and not args.no_cpu_baseline:
            g16 = line.pop("_gpu16")
            s = CpuSample()
            dt, cpu_proof = s.prove()
            line["cpu_baseline"] = {"value": s.n_constraints / dt, "unit": "constraints/s", "cores": s.cores, "kind": "port", "sample": s.describe(),
                                    "config": SHARED_DESC, "cpu_ms": dt * 1e3, "gpu_same_config_ms": g16["ms"],
                                    "gpu_same_config_constraints_per_s": g16["constraints"] / (g16["ms"] * 1e-3),
                                    "same_proof_bytes": hashlib.sha256(cpu_proof).hexdigest() == g16["proof_sha256"], "gpu_proof_verified": g16["verified"],
                                    "extrapolation": EXTRAPOLATION}
        emit(line)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
