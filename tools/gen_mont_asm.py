#!/usr/bin/env python3
"""Generator + bit-accurate simulator for the even/odd-split Montgomery multiplier
used by the sm_100a kernels (csrc/ff_mont_asm.inc).

The same row descriptions drive (a) a Python model of the PTX carry-flag semantics
(mad.lo.cc / madc.hi.cc / addc) that is checked here against big-int arithmetic and
(b) the emitted inline-PTX, so the CUDA text cannot drift from the verified model.

Representation during the multiply: two n-limb arrays E and O with
    T = sum E[k] 2^(32k) + sum O[k] 2^(32(k+1))
so 64-bit partial products of even-indexed a[j] land on (E[j],E[j+1]) and those of
odd-indexed a[j] on (O[j-1],O[j]); both are single carry chains.  After each
reduction row E[0]==0 and the roles of E and O swap (division by 2^32).
"""
import random, sys

MASK = 0xFFFFFFFF

class Sim:
    """PTX carry-flag model; every row is a list of (op, dst, a, b, c) tuples."""
    def __init__(self):
        self.cc = 0
    def run(self, row, env):
        for ins in row:
            op, d, x, y, z = ins
            g = lambda t: env[t] if isinstance(t, str) else t
            if op == "taint":   # d = d | (x & zero): a scheduling fence, value-preserving because zero == 0
                env[d] = g(d) | (g(x) & g(y))
                continue
            if op == "shl1":    # d = x << 1 (mod 2^32)
                env[d] = (g(x) << 1) & MASK
                continue
            if op == "shf1":    # d = funnel shift left by one of (y : x): (y << 1) | (x >> 31)
                env[d] = ((g(y) << 1) | (g(x) >> 31)) & MASK
                continue
            if op in ("mul.lo", "mul.hi"):
                p = g(x) * g(y)
                env[d] = (p & MASK) if op == "mul.lo" else (p >> 32)
                continue
            if op.startswith("mad"):
                p = g(x) * g(y)
                p = (p & MASK) if ".lo" in op else (p >> 32)
                cin = self.cc if op.startswith("madc") else 0
                s = p + g(z) + cin
            elif op.startswith("add"):
                cin = self.cc if op.startswith("addc") else 0
                s = g(x) + g(y) + cin
            else:
                raise ValueError(op)
            env[d] = s & MASK
            if op.endswith(".cc"):
                self.cc = s >> 32
            else:
                assert s >> 32 == 0 or op in ("add", "addc", "mad.lo", "madc.hi", "madc.lo"), op
                env["_ovf"] = env.get("_ovf", 0) | (s >> 32)

def rows_for(n, square=False):
    """Return the list of rows (each a list of instrs) for an n-limb Montgomery product.
    Symbols: a0..a{n-1}, b0.., p0.., m, inv, X0..X{n-1}/Y0.. accumulators, r0.. result.
    square=True: the rows of sqr_rows_for (multiplier a_i, multiplicand vector with skipped / doubled columns)."""
    assert n % 2 == 0
    rows = []
    E, O = "X", "Y"
    def cand(i, j):  # multiplicand of column j in row i (None = skipped)
        if not square:
            return f"a{j}"
        return None if j < i else f"a{j}" if j == i else f"e{j}" if j == i + 1 else f"d{j}"
    for i in range(n):
        bi = f"a{i}" if square else f"b{i}"
        if i == 0:
            r = []
            for j in range(0, n, 2):
                r.append(("mul.lo", f"{E}{j}", cand(i, j), bi, None))
                r.append(("mul.hi", f"{E}{j+1}", cand(i, j), bi, None))
            rows.append(r)
            r = []
            for j in range(1, n, 2):
                r.append(("mul.lo", f"{O}{j-1}", cand(i, j), bi, None))
                r.append(("mul.hi", f"{O}{j}", cand(i, j), bi, None))
            rows.append(r)
        else:
            # fence: make this row's head depend on the previous row's last carry-out (E[n-1] is where both of the
            # previous row's chains ended).  Without it ptxas hoists the heads of all n rows (they only depend on
            # each other), keeps >7 carry chains live at once and spills carries through P2R/LOP3/ISETP.
            rows.append([("taint", f"{E}0", f"{E}{n-1}", "zero", None)])
            # S1: fold O[1] (position 0 after the shift) into E[0]; O <- (O >> 2 limbs) + a_odd*bi
            r = [("add.cc", f"{E}0", f"{E}0", f"{O}1", None)]
            for j in range(1, n, 2):
                src_lo = f"{O}{j+1}" if j + 1 < n else 0
                src_hi = f"{O}{j+2}" if j + 2 < n else 0
                last = (j == n - 1)
                if cand(i, j) is None:  # skipped column: the two limbs still move down (O >> 2 limbs) and pass the carry on
                    r.append(("addc.cc", f"{O}{j-1}", src_lo, 0, None))
                    r.append(("addc" if last else "addc.cc", f"{O}{j}", src_hi, 0, None))
                else:
                    r.append(("madc.lo.cc", f"{O}{j-1}", cand(i, j), bi, src_lo))
                    r.append(("madc.hi" if last else "madc.hi.cc", f"{O}{j}", cand(i, j), bi, src_hi))
            rows.append(r)
            # S2: E += a_even*bi ; carry -> O[n-1]
            r = []
            for j in range(0, n, 2):
                if cand(i, j) is None:
                    continue  # skipped column below the first product: nothing to add, no carry can reach it
                r.append(("mad.lo.cc" if not r else "madc.lo.cc", f"{E}{j}", cand(i, j), bi, f"{E}{j}"))
                r.append(("madc.hi.cc", f"{E}{j+1}", cand(i, j), bi, f"{E}{j+1}"))
            if r:
                r.append(("addc", f"{O}{n-1}", f"{O}{n-1}", 0, None))
                rows.append(r)
        rows.append([("mul.lo", "m", f"{E}0", "inv", None)])
        # S3: O += p_odd*m (no carry out)
        r = []
        for j in range(1, n, 2):
            last = (j == n - 1)
            r.append(("mad.lo.cc" if j == 1 else "madc.lo.cc", f"{O}{j-1}", f"p{j}", "m", f"{O}{j-1}"))
            r.append(("madc.hi" if last else "madc.hi.cc", f"{O}{j}", f"p{j}", "m", f"{O}{j}"))
        rows.append(r)
        # S4: E += p_even*m ; carry -> O[n-1]; E[0] becomes 0
        r = []
        for j in range(0, n, 2):
            r.append(("mad.lo.cc" if j == 0 else "madc.lo.cc", f"{E}{j}", f"p{j}", "m", f"{E}{j}"))
            r.append(("madc.hi.cc", f"{E}{j+1}", f"p{j}", "m", f"{E}{j+1}"))
        r.append(("addc", f"{O}{n-1}", f"{O}{n-1}", 0, None))
        rows.append(r)
        E, O = O, E
    # after n (even) swaps E=="X", O=="Y" hold the *pre-shift* state with roles swapped back:
    # the array that just got its limb 0 cleared is O (because of the swap); T/B = E + (O>>1 limb)
    r = [("add.cc", "r0", f"{E}0", f"{O}1", None)]
    for k in range(1, n - 1):
        r.append(("addc.cc", f"r{k}", f"{E}{k}", f"{O}{k+1}", None))
    r.append(("addc", f"r{n-1}", f"{E}{n-1}", 0, None))
    rows.append(r)
    return rows

def simulate(n, a, b, p):
    inv = (-pow(p, -1, 1 << 32)) & MASK
    env = {"inv": inv, "zero": 0}
    for k in range(n):
        env[f"a{k}"] = (a >> (32 * k)) & MASK
        env[f"b{k}"] = (b >> (32 * k)) & MASK
        env[f"p{k}"] = (p >> (32 * k)) & MASK
    sim = Sim()
    for row in rows_for(n):
        sim.run(row, env)
    assert env.get("_ovf", 0) == 0, "unexpected overflow on a non-.cc op"
    r = sum(env[f"r{k}"] << (32 * k) for k in range(n))
    return r

def sqr_rows_for(n):
    """Rows of a dedicated Montgomery SQUARING r = a*a/2^(32n) mod p (result < 2p) in the SAME row structure as the product (fixed even/odd
    accumulator pairs, so ptxas needs no register re-pairing):  a^2 = sum_i a_i 2^(32i) * (a_i 2^(32i) + 2 U_i 2^(32(i+1))),  U_i = a >> 32(i+1).
    Row i multiplies a_i with the vector [ -, .., -, a_i, e_(i+1), d_(i+2), .., d_(n-1) ]: columns j < i are SKIPPED (their products were
    counted, doubled, in earlier rows), e_j = a_j << 1 and d_j = (a_j << 1) | (a_(j-1) >> 31) are the limbs of the doubled upper part
    (needs 3p < 2^(32n): the doubled multiplicand makes the running sum reach 3p; true for the two 12-limb base fields, NOT for Fr).  A skipped product keeps its place in the carry chain as
    a plain add.  n(n+1)/2 + n^2 wide multiplies instead of 2 n^2: 222 instead of 288 for n = 12."""
    pre = []
    for j in range(1, n):
        pre.append(("shl1", f"e{j}", f"a{j}", None, None))
        pre.append(("shf1", f"d{j}", f"a{j-1}", f"a{j}", None))
    return [pre] + rows_for(n, square=True)

def simulate_sqr(n, a, p):
    inv = (-pow(p, -1, 1 << 32)) & MASK
    env = {"inv": inv, "zero": 0}
    for k in range(n):
        env[f"a{k}"] = (a >> (32 * k)) & MASK
        env[f"p{k}"] = (p >> (32 * k)) & MASK
    sim = Sim()
    for row in sqr_rows_for(n):
        sim.run(row, env)
    assert env.get("_ovf", 0) == 0, "unexpected overflow on a non-.cc op"
    return sum(env[f"r{k}"] << (32 * k) for k in range(n))

def selftest():
    P377 = 0x1ae3a4617c510eac63b05c06ca1493b1a22d9f300f5138f1ef3622fba094800170b5d44300000008508c00000000001
    R377 = 0x12ab655e9a2ca55660b44d1e5c37b00159aa76fed00000010a11800000000001
    P381 = 0x1a0111ea397fe69a4b1ba7b6434bacd764774b84f38512bf6730d2a0f6b0f6241eabfffeb153ffffb9feffffffffaaab
    R381 = 0x73eda753299d7d483339d80809a1d80553bda402fffe5bfeffffffff00000001
    rnd = random.Random(1)
    for n, p in ((12, P377), (8, R377), (12, P381), (8, R381)):
        Rinv = pow(1 << (32 * n), -1, p)
        cases = [(0, 0), (p - 1, p - 1), (1, p - 1), (p - 1, 1)]
        cases += [(rnd.randrange(p), rnd.randrange(p)) for _ in range(300)]
        for a, b in cases:
            r = simulate(n, a, b, p)
            assert r < 2 * p, "result not < 2p"
            assert r % p == a * b * Rinv % p, (n, hex(a), hex(b))
        if n != 12:
            continue  # the dedicated squaring needs 3p < 2^(32n) (doubled multiplicand): true for the two 12-limb base fields, not for Fr
        allones = (1 << (32 * n)) - 1
        for a in [0, 1, p - 1, p - 2, p >> 1, allones % p, (allones >> 7) % p, sum(0xffffffff << (64 * k) for k in range(n // 2)) % p] + [x for x, _ in cases[4:]]:
            r = simulate_sqr(n, a, p)
            assert r < 2 * p, "square not < 2p"
            assert r % p == a * a * Rinv % p, (n, hex(a))
    print("gen_mont_asm: model OK (n=8,12; both curves; products and dedicated squarings)")

def emit(n, name, square=False):
    """Emit a __device__ function whose body is one asm statement per row."""
    out = []
    out.append(f"// GENERATED by tools/gen_mont_asm.py -- do not edit. n={n} 32-bit limbs.")
    if square:
        out.append(f"// Dedicated squaring (sqr_rows_for): r = a*a/2^{32*n} mod p, result in [0, 2p); {n*(n+1)//2 + n*n} wide multiplies instead of {2*n*n}.")
        out.append(f"__device__ __forceinline__ void {name}(uint32_t* __restrict__ r, const uint32_t* __restrict__ a,")
        out.append(f"        const uint32_t* __restrict__ p, uint32_t inv, uint32_t zero) {{")
        out.append(f"    uint32_t X[{n}], Y[{n}], e[{n}], d[{n}], m;")
    else:
        out.append(f"// r = a*b/2^{32*n} mod p, result in [0, 2p); caller does the final conditional subtract.")
        out.append(f"__device__ __forceinline__ void {name}(uint32_t* __restrict__ r, const uint32_t* __restrict__ a,")
        out.append(f"        const uint32_t* __restrict__ b, const uint32_t* __restrict__ p, uint32_t inv, uint32_t zero) {{")
        out.append(f"    uint32_t X[{n}], Y[{n}], m;")
    def ref(t, ops, kinds):
        # map symbol -> %k placeholder, registering operand
        if not isinstance(t, str):
            return str(t)
        if t not in ops:
            ops[t] = len(ops)
        return None
    all_rows = sqr_rows_for(n) if square else rows_for(n)
    for row in all_rows[:-1]:
        # collect symbols: written ones first
        written, read = [], []
        for op, d, x, y, z in row:
            if d not in written:
                written.append(d)
            for t in ((d, x, y) if op == "taint" else (x, y, z)):
                if isinstance(t, str) and t not in read:
                    read.append(t)
        # an operand that is written is "+r" if it is read before/at all, else "=r"
        # conservative: "+r" whenever it also appears in read, "=r" otherwise -- but "=r"
        # operands written early and read later in the same asm need early-clobber; use "+r"
        # for everything written except pure outputs never read in this row ("=&r").
        order, cons = [], []
        def cexpr(t):
            if t == "m": return "m"
            if t == "inv": return "inv"
            if t == "zero": return "zero"
            arr, idx = t[0], int(t[1:])
            return {"X": "X", "Y": "Y", "a": "a", "b": "b", "p": "p", "r": "r", "e": "e", "d": "d"}[arr] + f"[{idx}]"
        # first-use analysis for written symbols
        first_is_write = {}
        for op, d, x, y, z in row:
            for t in ((d, x, y) if op == "taint" else (x, y, z)):
                if isinstance(t, str) and t not in first_is_write:
                    first_is_write[t] = False
            if d not in first_is_write:
                first_is_write[d] = True
        outs = [(t, "=&r" if first_is_write[t] else "+r") for t in written]
        ins = [t for t in read if t not in written]
        idx = {}
        for t, _ in outs:
            idx[t] = len(idx)
        for t in ins:
            idx[t] = len(idx)
        def o(t):
            return f"%{idx[t]}" if isinstance(t, str) else str(t)
        lines = []
        for op, d, x, y, z in row:
            if op == "taint":
                lines.append(f"lop3.b32 {o(d)}, {o(d)}, {o(x)}, {o(y)}, 0xF8;")
            elif op == "shl1":
                lines.append(f"shl.b32 {o(d)}, {o(x)}, 1;")
            elif op == "shf1":
                lines.append(f"shf.l.wrap.b32 {o(d)}, {o(x)}, {o(y)}, 1;")
            elif op in ("mul.lo", "mul.hi"):
                lines.append(f"{op}.u32 {o(d)}, {o(x)}, {o(y)};")
            elif op.startswith("mad"):
                lines.append(f"{op}.u32 {o(d)}, {o(x)}, {o(y)}, {o(z)};")
            else:
                lines.append(f"{op}.u32 {o(d)}, {o(x)}, {o(y)};")
        body = " ".join(lines)
        outc = ", ".join(f'"{c}"({cexpr(t)})' for t, c in outs)
        inc = ", ".join(f'"r"({cexpr(t)})' for t in ins)
        out.append(f'    asm("{body}"\n        : {outc}\n        : {inc});')
    # final row (T/B = X + (Y >> one limb)) in plain C++: ptxas emits an IADD3 carry chain
    out.append("    uint64_t c = 0;")
    out.append(f"    #pragma unroll")
    out.append(f"    for (int k = 0; k < {n}; ++k) {{")
    out.append(f"        c += (uint64_t)X[k] + (k + 1 < {n} ? Y[k + 1] : 0u);")
    out.append("        r[k] = (uint32_t)c; c >>= 32;")
    out.append("    }")
    out.append("}")
    return "\n".join(out)

if __name__ == "__main__":
    selftest()
    if len(sys.argv) > 1:
        with open(sys.argv[1], "w") as f:
            f.write("// clang-format off\n")
            f.write(emit(8, "mont_mul_raw_8") + "\n\n")
            f.write(emit(12, "mont_mul_raw_12") + "\n\n")
            f.write(emit(12, "mont_sqr_raw_12", square=True) + "\n")
        print("wrote", sys.argv[1])
