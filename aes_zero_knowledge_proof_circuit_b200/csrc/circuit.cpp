// Host-side synthesis of the AES-128-ECB R1CS (see circuit.h).  Plain C++17, no CUDA.
#include "circuit.h"

#include <algorithm>
#include <stdexcept>
#include <string>

namespace zk {
namespace {

// ---- variables: ark-relations 0.3.0 `Variable` restricted to One / Instance(i) / Witness(i) ----------------------
using Var = uint32_t;
constexpr Var VAR_ONE = 0;
constexpr Var TAG_INST = 1u << 30, TAG_WIT = 2u << 30, TAG_MASK = 3u << 30;

// ---- ark-r1cs-std 0.3.1 Boolean: Constant(bool) | Is(var) | Not(var) ---------------------------------------------
struct Bool {
    uint8_t kind;  // 0 const, 1 is, 2 not
    uint8_t cval;  // for constants
    Var var;
    bool is_const() const { return kind == 0; }
    bool is_true() const { return kind == 0 && cval; }
    bool is_false() const { return kind == 0 && !cval; }
};
constexpr Bool B_TRUE{0, 1, 0}, B_FALSE{0, 0, 0};
inline Bool b_not(Bool b) {
    if (b.kind == 0) return Bool{0, (uint8_t)!b.cval, 0};
    return Bool{(uint8_t)(b.kind == 1 ? 2 : 1), 0, b.var};
}
using Byte = Bool[8];
struct U8 {
    Bool b[8];  // LSB first
};

struct Term {
    int32_t coeff;
    Var var;
};

struct Builder {
    AesCircuit& out;
    uint32_t n_inst = 1;  // Variable::One is instance 0
    uint32_t n_wit = 0;
    uint32_t inst_padded;
    // witness-program recording
    enum Mode { REC_NONE, REC_FIXED, REC_BLOCK } mode = REC_NONE;
    uint32_t cur_block = 0, cur_block_start = 0, cur_instr = 0;
    std::vector<WitInstr> fixed_instrs, block_instrs;
    size_t msg_bits;

    explicit Builder(AesCircuit& o, size_t msg_len) : out(o), msg_bits(8 * msg_len) {
        uint32_t used = 1 + (uint32_t)msg_bits;
        inst_padded = 1;
        while (inst_padded < used) inst_padded <<= 1;
        for (CsrMatrix* m : {&out.a, &out.b, &out.c}) m->row_ptr.push_back(0);
    }

    uint32_t col_of(Var v) const {
        switch (v & TAG_MASK) {
            case 0: return 0;
            case TAG_INST: return v & ~TAG_MASK;
            default: return inst_padded + (v & ~TAG_MASK);
        }
    }
    Var new_witness() { return TAG_WIT | n_wit++; }
    Var new_input() { return TAG_INST | n_inst++; }

    // LinearCombination: sorted by variable, equal variables merged, zero coefficients dropped by to_matrices
    void push_row(CsrMatrix& m, const Term* t, int n) {
        Term tmp[8];
        int k = 0;
        for (int i = 0; i < n; ++i) {
            int j = 0;
            for (; j < k; ++j)
                if (tmp[j].var == t[i].var) break;
            if (j < k)
                tmp[j].coeff += t[i].coeff;
            else
                tmp[k++] = t[i];
        }
        std::pair<uint32_t, int32_t> e[8];
        int ne = 0;
        for (int i = 0; i < k; ++i)
            if (tmp[i].coeff != 0) e[ne++] = {col_of(tmp[i].var), tmp[i].coeff};
        for (int i = 1; i < ne; ++i)  // at most 8 entries: insertion sort by column (std::sort's 16-element unrolling trips -Warray-bounds)
            for (int j = i; j > 0 && e[j] < e[j - 1]; --j) std::swap(e[j], e[j - 1]);
        for (int i = 0; i < ne; ++i) {
            if (e[i].second < -128 || e[i].second > 127) throw std::runtime_error("circuit: coefficient out of int8 range");
            m.col.push_back(e[i].first);
            m.coeff.push_back((int8_t)e[i].second);
        }
        m.row_ptr.push_back((uint32_t)m.col.size());
    }
    void enforce(const Term* a, int na, const Term* b, int nb, const Term* c, int nc) {
        push_row(out.a, a, na);
        push_row(out.b, b, nb);
        push_row(out.c, c, nc);
    }
    // Boolean::lc(): Is(v) -> v ; Not(v) -> 1 - v ; Constant(true) -> 1 ; Constant(false) -> 0
    static int lc(Bool b, Term* t, int sign = 1) {
        if (b.kind == 1) {
            t[0] = {sign, b.var};
            return 1;
        }
        if (b.kind == 2) {
            t[0] = {sign, VAR_ONE};
            t[1] = {-sign, b.var};
            return 2;
        }
        if (b.cval) {
            t[0] = {sign, VAR_ONE};
            return 1;
        }
        return 0;
    }

    // ---- witness program -----------------------------------------------------------------------------------------
    uint32_t ref_of(Bool b) const {
        if (b.kind == 0) return REF_CONST | b.cval;
        uint32_t neg = b.kind == 2 ? REF_NEG : 0;
        if ((b.var & TAG_MASK) != TAG_WIT) throw std::runtime_error("circuit: gadget operand is not a witness");
        uint32_t i = b.var & ~TAG_MASK;
        if (mode == REC_BLOCK) {
            if (i < msg_bits) {
                uint32_t lo = 128 * cur_block;
                if (i < lo || i >= lo + 128) throw std::runtime_error("circuit: block reads another block's message bits");
                return REF_MSG | neg | (i - lo);
            }
            if (i >= cur_block_start) return REF_LOCAL | neg | (i - cur_block_start);
            if (i >= out.wit_fixed_end) throw std::runtime_error("circuit: block reads another block's witnesses");
        }
        return REF_GLOBAL | neg | i;
    }
    void record(uint32_t op, Var dst, uint32_t a, uint32_t b, uint32_t c) {
        uint32_t d = dst & ~TAG_MASK;
        if (mode == REC_FIXED) {
            fixed_instrs.push_back({d, a, b, c, op});
        } else if (mode == REC_BLOCK) {
            WitInstr ins{d - cur_block_start, a, b, c, op};
            if (cur_block == 0) {
                block_instrs.push_back(ins);
            } else {  // every block must repeat block 0's program exactly (that is what lets the GPU reuse it)
                if (cur_instr >= block_instrs.size()) throw std::runtime_error("circuit: block program length differs");
                const WitInstr& r = block_instrs[cur_instr];
                if (r.dst != ins.dst || r.a != ins.a || r.b != ins.b || r.c != ins.c || r.op != ins.op)
                    throw std::runtime_error("circuit: block " + std::to_string(cur_block) + " is not a shifted copy of block 0");
            }
            ++cur_instr;
        } else {
            throw std::runtime_error("circuit: gadget witness allocated outside a recorded section");
        }
    }

    // ---- gadgets ---------------------------------------------------------------------------------------------------
    Bool alloc_bool(bool input) {  // AllocatedBool::new_variable: (1 - b) * b = 0
        Var v = input ? new_input() : new_witness();
        Term a[2] = {{1, VAR_ONE}, {-1, v}}, b[1] = {{1, v}};
        enforce(a, 2, b, 1, nullptr, 0);
        return Bool{1, 0, v};
    }
    Bool bxor(Bool a, Bool b) {
        if (a.is_false()) return b;
        if (b.is_false()) return a;
        if (a.is_true()) return b_not(b);
        if (b.is_true()) return b_not(a);
        if (a.kind != b.kind) {
            Bool is = a.kind == 1 ? a : b, nt = a.kind == 1 ? b : a;
            return b_not(bxor(is, b_not(nt)));
        }
        Var r = new_witness();
        Term ta[2] = {{1, a.var}, {1, a.var}}, tb[1] = {{1, b.var}}, tc[3] = {{1, a.var}, {1, b.var}, {-1, r}};
        enforce(ta, 2, tb, 1, tc, 3);
        record(WOP_XOR, r, ref_of(Bool{1, 0, a.var}), ref_of(Bool{1, 0, b.var}), 0);
        return Bool{1, 0, r};
    }
    Bool band(Bool a, Bool b) {
        if (a.is_false() || b.is_false()) return B_FALSE;
        if (a.is_true()) return b;
        if (b.is_true()) return a;
        Var r = new_witness();
        Term ta[2], tb[2], tc[1] = {{1, r}};
        int na, nb;
        if (a.kind == 1 && b.kind == 1) {
            na = lc(a, ta);
            nb = lc(b, tb);
        } else if (a.kind == 2 && b.kind == 2) {  // nor
            na = lc(a, ta);
            nb = lc(b, tb);
        } else {  // and_not: (is) * (1 - not)
            Bool is = a.kind == 1 ? a : b, nt = a.kind == 1 ? b : a;
            na = lc(is, ta);
            nb = lc(nt, tb);
        }
        enforce(ta, na, tb, nb, tc, 1);
        record(WOP_AND, r, ref_of(a), ref_of(b), 0);
        return Bool{1, 0, r};
    }
    // ark-r1cs-std 0.3.1 Boolean::or: constants fold; (Is, Is) is AllocatedBool::or -- a fresh witness r = a | b with
    // (1 - a) * (1 - b) = (1 - r), result Is(r); every other combination is NOT(AND(NOT a, NOT b)).
    Bool bor(Bool a, Bool b) {
        if (a.is_false()) return b;
        if (b.is_false()) return a;
        if (a.is_true() || b.is_true()) return B_TRUE;
        if (a.kind == 1 && b.kind == 1) {
            Var r = new_witness();
            Term ta[2] = {{1, VAR_ONE}, {-1, a.var}}, tb[2] = {{1, VAR_ONE}, {-1, b.var}}, tc[2] = {{1, VAR_ONE}, {-1, r}};
            enforce(ta, 2, tb, 2, tc, 2);
            record(WOP_SEL, r, ref_of(a), REF_CONST | 1u, ref_of(b));  // a ? 1 : b
            return Bool{1, 0, r};
        }
        return b_not(band(b_not(a), b_not(b)));
    }
    Bool select(Bool cond, Bool t, Bool f) {
        if (cond.is_true()) return t;
        if (cond.is_false()) return f;
        if (cond.kind == 2) return select(b_not(cond), f, t);
        if (f.is_false()) return band(cond, t);
        if (t.is_false()) return band(b_not(cond), f);
        if (t.is_true()) return bor(cond, f);
        if (f.is_true()) return bor(b_not(cond), t);
        Var r = new_witness();
        Term ta[2], tb[4], tc[3];
        int na = lc(cond, ta);
        int nb = lc(t, tb);
        nb += lc(f, tb + nb, -1);
        tc[0] = {1, r};
        int nc = 1 + lc(f, tc + 1, -1);
        enforce(ta, na, tb, nb, tc, nc);
        record(WOP_SEL, r, ref_of(cond), ref_of(t), ref_of(f));
        return Bool{1, 0, r};
    }
    void enforce_equal(Bool a, Bool b) {  // Boolean::conditional_enforce_equal with condition TRUE
        Term d[3];
        int nd = 0;
        if (a.is_const() && b.is_const()) {
            if (a.cval != b.cval) throw std::runtime_error("circuit: unsatisfiable constant equality");
            return;
        }
        if (a.is_const() || b.is_const()) {
            Bool c = a.is_const() ? a : b, x = a.is_const() ? b : a;
            bool one_minus = c.cval ? (x.kind == 1) : (x.kind == 2);
            if (one_minus) {
                d[nd++] = {1, VAR_ONE};
                d[nd++] = {-1, x.var};
            } else {
                d[nd++] = {1, x.var};
            }
        } else if (a.kind == 1 && b.kind == 1) {
            d[nd++] = {1, b.var};
            d[nd++] = {-1, a.var};
        } else if (a.kind == 2 && b.kind == 2) {
            d[nd++] = {1, a.var};
            d[nd++] = {-1, b.var};
        } else {
            Bool is = a.kind == 1 ? a : b, nt = a.kind == 1 ? b : a;
            d[nd++] = {1, VAR_ONE};
            d[nd++] = {-1, nt.var};
            d[nd++] = {-1, is.var};
        }
        Term one[1] = {{1, VAR_ONE}};
        enforce(d, nd, one, 1, nullptr, 0);
    }

    // ---- UInt8 -------------------------------------------------------------------------------------------------------
    static U8 const_byte(uint8_t v) {
        U8 r;
        for (int i = 0; i < 8; ++i) r.b[i] = ((v >> i) & 1) ? B_TRUE : B_FALSE;
        return r;
    }
    U8 new_byte(bool input) {
        U8 r;
        for (int i = 0; i < 8; ++i) r.b[i] = alloc_bool(input);
        return r;
    }
    U8 xor8(const U8& a, const U8& b) {
        U8 r;
        for (int i = 0; i < 8; ++i) r.b[i] = bxor(a.b[i], b.b[i]);
        return r;
    }
    U8 select8(Bool c, const U8& t, const U8& f) {
        U8 r;
        for (int i = 0; i < 8; ++i) r.b[i] = select(c, t.b[i], f.b[i]);
        return r;
    }
};

uint8_t g_sbox[256];
bool g_sbox_ready = false;
void init_sbox() {  // src/aes.rs:24-62
    if (g_sbox_ready) return;
    auto rotl8 = [](uint8_t x, int s) { return (uint8_t)((x << s) | (x >> (8 - s))); };
    uint8_t p = 1, q = 1;
    do {
        p = (uint8_t)(p ^ (p << 1) ^ ((p & 0x80) ? 0x1B : 0));
        q ^= (uint8_t)(q << 1);
        q ^= (uint8_t)(q << 2);
        q ^= (uint8_t)(q << 4);
        if (q & 0x80) q ^= 0x09;
        uint8_t x = (uint8_t)(q ^ rotl8(q, 1) ^ rotl8(q, 2) ^ rotl8(q, 3) ^ rotl8(q, 4));
        g_sbox[p] = (uint8_t)(x ^ 0x63);
    } while (p != 1);
    g_sbox[0] = 0x63;
    g_sbox_ready = true;
}

// src/aes_circuit.rs:243-248: UInt8::conditionally_select_power_of_two_vector(byte.to_bits_be(), table)
U8 sub_byte(Builder& B, const U8& in) {
    std::vector<U8> cur(256);
    for (int v = 0; v < 256; ++v) cur[v] = Builder::const_byte(g_sbox[v]);
    // position = to_bits_be(): position[n-1-i] is bit i (LSB first level)
    for (int i = 0; i < 8; ++i) {
        std::vector<U8> nxt(cur.size() / 2);
        for (size_t j = 0; j < cur.size(); j += 2) nxt[j / 2] = B.select8(in.b[i], cur[j + 1], cur[j]);
        cur.swap(nxt);
    }
    return cur[0];
}
// simpleworks shift gadgets: modelled as rewiring with constant-false fill
U8 shl(const U8& a, int k) {
    U8 r;
    for (int i = 0; i < 8; ++i) r.b[i] = i < k ? B_FALSE : a.b[i - k];
    return r;
}
U8 shr(const U8& a, int k) {
    U8 r;
    for (int i = 0; i < 8; ++i) r.b[i] = i + k < 8 ? a.b[i + k] : B_FALSE;
    return r;
}
// src/helpers/mod.rs:11-42
U8 add8(Builder& B, const U8& a, const U8& b) {
    U8 s;
    Bool carry = B_FALSE;
    for (int i = 0; i < 8; ++i) {  // big-endian index 7-i == little-endian bit i, processed LSB first
        s.b[i] = B.bxor(B.bxor(carry, a.b[i]), b.b[i]);
        // gadget calls in the reference's evaluation order (each may allocate a witness)
        Bool ab = B.band(a.b[i], b.b[i]);
        Bool gen = B.band(b_not(carry), ab);
        Bool aob = B.bor(a.b[i], b.b[i]);
        Bool prop = B.band(carry, aob);
        carry = B.bor(gen, prop);
    }
    return s;
}
// src/helpers/mod.rs:44-64 with a constant multiplier (its VALUE drives the branches)
U8 mul_const(Builder& B, const U8& h, uint8_t m) {
    U8 prod = Builder::const_byte(0);
    for (int i = 0; i < 8; ++i)
        if ((m >> i) & 1) prod = add8(B, prod, i ? shl(h, i) : h);
    return prod;
}
// src/aes_circuit.rs:360-427
void gmix_column(Builder& B, const U8* col, U8* o) {
    U8 b[4];
    const U8 one = Builder::const_byte(1);
    for (int k = 0; k < 4; ++k) {
        U8 sr = shr(col[k], 7), h;
        for (int i = 0; i < 8; ++i) h.b[i] = B.band(sr.b[i], one.b[i]);
        b[k] = B.xor8(shl(col[k], 1), mul_const(B, h, 0x1B));
    }
    auto X5 = [&](const U8& p, const U8& q, const U8& r, const U8& s, const U8& t) { return B.xor8(B.xor8(B.xor8(B.xor8(p, q), r), s), t); };
    o[0] = X5(b[0], col[3], col[2], b[1], col[1]);
    o[1] = X5(b[1], col[0], col[3], b[2], col[2]);
    o[2] = X5(b[2], col[1], col[0], b[3], col[3]);
    o[3] = X5(b[3], col[2], col[1], b[0], col[0]);
}
void mix_columns(Builder& B, U8* st) {
    U8 o[16];
    for (int i = 0; i < 4; ++i) gmix_column(B, st + 4 * i, o + 4 * i);
    std::copy(o, o + 16, st);
}
// src/aes_circuit.rs:268-334 (rotate_left modelled as rewiring)
void shift_rows(U8* st) {
    static const int map[16] = {0, 5, 10, 15, 4, 9, 14, 3, 8, 13, 2, 7, 12, 1, 6, 11};
    U8 o[16];
    for (int i = 0; i < 16; ++i) o[i] = st[map[i]];
    std::copy(o, o + 16, st);
}
void add_round_key(Builder& B, U8* st, const U8* rk) {
    for (int i = 0; i < 16; ++i) st[i] = B.xor8(st[i], rk[i]);
}
struct U32 {
    Bool b[32];  // LSB first
};
U32 to_u32(const U8* be) {  // src/aes_circuit.rs:200-212
    U32 r;
    for (int k = 0; k < 4; ++k)
        for (int i = 0; i < 8; ++i) r.b[8 * k + i] = be[3 - k].b[i];
    return r;
}
void to_bytes_be(const U32& w, U8* out) {  // src/aes_circuit.rs:188-198
    for (int k = 0; k < 4; ++k)
        for (int i = 0; i < 8; ++i) out[k].b[i] = w.b[8 * (3 - k) + i];
}
// src/aes_circuit.rs:20-129
void derive_keys(Builder& B, const U8* key, U8 rks[11][16]) {
    static const uint8_t rc[10] = {0x01, 0x02, 0x04, 0x08, 0x10, 0x20, 0x40, 0x80, 0x1B, 0x36};
    std::vector<U32> W(44);
    for (int i = 0; i < 4; ++i) W[i] = to_u32(key + 4 * i);
    for (int i = 4; i < 44; ++i) {
        if (i % 4 == 0) {
            U8 wb[4], rot[4], sub[4];
            to_bytes_be(W[i - 1], wb);
            for (int k = 0; k < 4; ++k) rot[k] = wb[(k + 1) % 4];
            for (int k = 0; k < 4; ++k) sub[k] = sub_byte(B, rot[k]);
            U32 sw = to_u32(sub);
            U8 rcb[4] = {Builder::const_byte(rc[i / 4 - 1]), Builder::const_byte(0), Builder::const_byte(0), Builder::const_byte(0)};
            U32 rcw = to_u32(rcb);
            U32 res;
            for (int j = 0; j < 32; ++j) res.b[j] = B.bxor(W[i - 4].b[j], sw.b[j]);
            for (int j = 0; j < 32; ++j) res.b[j] = B.bxor(res.b[j], rcw.b[j]);
            W[i] = res;
        } else {
            for (int j = 0; j < 32; ++j) W[i].b[j] = B.bxor(W[i - 4].b[j], W[i - 1].b[j]);
        }
    }
    for (int r = 0; r < 11; ++r)
        for (int k = 0; k < 4; ++k) to_bytes_be(W[4 * r + k], rks[r] + 4 * k);
}

// level = 1 + max(level of operands produced by the same program); stable sort by level
void levelize(std::vector<WitInstr>& instrs, uint32_t space, uint32_t base, uint32_t span, WitProgram& out) {
    std::vector<uint32_t> lvl_of(span, 0);  // level of each value in the program's own index space (0 = primary input)
    std::vector<uint32_t> lvl(instrs.size());
    uint32_t max_lvl = 0;
    auto operand_level = [&](uint32_t ref) -> uint32_t {
        if ((ref & (3u << 30)) != space) return 0;
        uint32_t i = (ref & REF_INDEX_MASK);
        if (i < base || i - base >= span) return 0;
        return lvl_of[i - base];
    };
    for (size_t k = 0; k < instrs.size(); ++k) {
        const WitInstr& in = instrs[k];
        uint32_t l = std::max(operand_level(in.a), operand_level(in.b));
        if (in.op == WOP_SEL) l = std::max(l, operand_level(in.c));
        lvl[k] = l + 1;
        lvl_of[in.dst - base] = l + 1;
        max_lvl = std::max(max_lvl, l + 1);
    }
    std::vector<uint32_t> count(max_lvl + 2, 0);
    for (uint32_t l : lvl) count[l + 1]++;
    for (size_t i = 1; i < count.size(); ++i) count[i] += count[i - 1];
    out.instrs.resize(instrs.size());
    std::vector<uint32_t> pos(count.begin(), count.end() - 1);
    for (size_t k = 0; k < instrs.size(); ++k) out.instrs[pos[lvl[k]]++] = instrs[k];
    // level_start over levels 1..max_lvl
    out.level_start.assign(count.begin() + 1, count.end());
}

}  // namespace

void build_aes_circuit(size_t msg_len, AesCircuit& out) {
    if (msg_len == 0 || msg_len % 16) throw std::runtime_error("message length must be a non-zero multiple of 16 bytes");
    if (msg_len > ((size_t)1 << 20)) throw std::runtime_error("message too long");
    init_sbox();
    out = AesCircuit();
    out.msg_len = msg_len;
    out.n_blocks = msg_len / 16;
    Builder B(out, msg_len);
    out.num_instance = B.inst_padded;
    out.num_instance_used = 1 + 8 * (uint32_t)msg_len;
    // src/lib.rs:70-88: message bytes, then key bytes, as witnesses
    std::vector<U8> msg(msg_len);
    for (size_t i = 0; i < msg_len; ++i) msg[i] = B.new_byte(false);
    out.wit_key0 = B.n_wit;
    U8 key[16];
    for (int i = 0; i < 16; ++i) key[i] = B.new_byte(false);
    out.wit_fixed0 = B.n_wit;
    // src/lib.rs:187: key schedule, once
    B.mode = Builder::REC_FIXED;
    U8 rks[11][16];
    derive_keys(B, key, rks);
    out.wit_fixed_end = B.n_wit;
    out.wit_block0 = B.n_wit;
    // src/lib.rs:194-277
    std::vector<U8> ct(msg_len);
    B.mode = Builder::REC_BLOCK;
    for (size_t blk = 0; blk < out.n_blocks; ++blk) {
        B.cur_block = (uint32_t)blk;
        B.cur_block_start = B.n_wit;
        B.cur_instr = 0;
        U8 st[16];
        std::copy(msg.begin() + 16 * blk, msg.begin() + 16 * blk + 16, st);
        add_round_key(B, st, key);  // round 0 uses the raw key
        for (int r = 1; r <= 9; ++r) {
            for (int i = 0; i < 16; ++i) st[i] = sub_byte(B, st[i]);
            shift_rows(st);
            mix_columns(B, st);
            add_round_key(B, st, rks[r]);
        }
        for (int i = 0; i < 16; ++i) st[i] = sub_byte(B, st[i]);
        shift_rows(st);
        add_round_key(B, st, rks[10]);
        if (blk == 0) {
            out.wit_block_stride = B.n_wit - B.cur_block_start;
            out.ct_refs.resize(128);
            for (int i = 0; i < 16; ++i)
                for (int j = 0; j < 8; ++j) out.ct_refs[8 * i + j] = B.ref_of(st[i].b[j]);
        } else {
            if (B.n_wit - B.cur_block_start != out.wit_block_stride || B.cur_instr != B.block_instrs.size())
                throw std::runtime_error("circuit: block size differs from block 0");
            for (int i = 0; i < 16; ++i)
                for (int j = 0; j < 8; ++j)
                    if (out.ct_refs[8 * i + j] != B.ref_of(st[i].b[j])) throw std::runtime_error("circuit: ciphertext wiring differs");
        }
        std::copy(st, st + 16, ct.begin() + 16 * blk);
    }
    B.mode = Builder::REC_NONE;
    // src/lib.rs:282-286: ciphertext as public input, byte by byte
    for (size_t i = 0; i < msg_len; ++i) {
        U8 p = B.new_byte(true);
        for (int j = 0; j < 8; ++j) B.enforce_equal(p.b[j], ct[i].b[j]);
    }
    if (B.n_inst != out.num_instance_used) throw std::runtime_error("circuit: instance count mismatch");
    out.num_witness_real = B.n_wit;
    // ark-marlin 0.3.0 constraint_systems.rs: make_matrices_square (instance already counted as padded)
    uint32_t ncons = (uint32_t)out.a.row_ptr.size() - 1;
    uint32_t nvar = out.num_instance + B.n_wit;
    if (nvar > ncons) {
        for (uint32_t i = ncons; i < nvar; ++i)
            for (CsrMatrix* m : {&out.a, &out.b, &out.c}) m->row_ptr.push_back((uint32_t)m->col.size());
        ncons = nvar;
    } else {
        B.n_wit += ncons - nvar;  // dummy witnesses with value one
    }
    out.num_witness = B.n_wit;
    out.num_constraints = ncons;
    levelize(B.fixed_instrs, REF_GLOBAL, out.wit_fixed0, out.wit_fixed_end - out.wit_fixed0, out.fixed_prog);
    levelize(B.block_instrs, REF_LOCAL, 0, out.wit_block_stride, out.block_prog);
}

}  // namespace zk
