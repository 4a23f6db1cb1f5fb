# round 2, GPU job 19 (1 GPU): memcheck over the batched fixed-base normalisation (SRS powers, n not a multiple of the run length) and a 16-byte key + proof
set -x
mkdir -p gpurun_out
timeout 110 compute-sanitizer --tool memcheck --error-exitcode 9 python - > gpurun_out/j19_memcheck_fb.log 2>&1 <<'PY'
import numpy as np, torch
import aes_zero_knowledge_proof_circuit_b200 as zk
ctx = zk.Context(0)
n = 100_003
d = torch.empty(n * 96, dtype=torch.uint8, device="cuda")
ctx.srs_powers_device(377, bytes(range(32)), n, d)
ctx.sync()
b = d.cpu().numpy().view(np.uint64).reshape(n, 12)
print("srs ok", int(b[-1, 0] != 0))
pk = ctx.synthesize_keys(16, bytes(range(32)), bytes(range(1, 33)))
ct, proof = ctx.encrypt(pk, bytes([1] * 16), bytes(16), bytes(range(32)))
print("proof", len(proof), zk.verify_encryption(pk.verifying_key(), proof, ct), pk.info["lagrange_points"])
PY
echo "memcheck rc=$?" >> gpurun_out/j19_memcheck_fb.log
tail -6 gpurun_out/j19_memcheck_fb.log
