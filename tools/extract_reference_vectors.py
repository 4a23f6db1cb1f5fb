#!/usr/bin/env python3
"""Extract the known-answer vectors the reference's own tests hold for the hot path into tests/golden/.

Reads (read-only) the reference checkout and writes small JSON fixtures, because /root/reference does not
exist on the GPU box.  Sources:
  * tests/integration_tests.rs:50-310  -- FIPS-197 Appendix B round trace (start / SubBytes / ShiftRows / MixColumns)
  * tests/integration_tests.rs:313-372 -- end-to-end plaintext/key/ciphertext/wrong-ciphertext (16 B and 64 B)
  * src/aes_circuit.rs:704-846, src/aes.rs:277-360 -- per-step gadget vectors (ARK, MixColumns, SubBytes, key expansion)
  * src/main.rs:11-12 -- the example input (message 01x16, key 00x16); its ciphertext is not in the reference and is
    recomputed here with the `cryptography` package.
Usage: python tools/extract_reference_vectors.py [/root/reference] [tests/golden]
"""
import json, os, re, sys

def arrays_after(src, name):
    """All [..] groups of hex bytes inside `let <name> = [ ... ];` (nested or flat)."""
    m = re.search(r"let\s+%s(?:\s*:\s*[^=]+)?\s*=\s*" % re.escape(name), src)
    assert m, name
    i = src.index("[", m.end())
    depth, j = 0, i
    while True:
        if src[j] == "[": depth += 1
        elif src[j] == "]":
            depth -= 1
            if depth == 0: break
        j += 1
    body = src[i:j + 1]
    body = re.sub(r"//[^\n]*", "", body)
    inner = re.findall(r"\[([^\[\]]*)\]", body)
    out = []
    for grp in inner:
        vals = re.findall(r"0x([0-9a-fA-F]{1,2})", grp)
        if vals: out.append(bytes(int(v, 16) for v in vals).hex())
    return out

def all_arrays_named(src, name):
    res = []
    for m in re.finditer(r"let\s+%s(?:\s*:\s*[^=]+)?\s*=\s*" % re.escape(name), src):
        res.append(arrays_after(src[m.start():], name))
    return res

def main():
    ref = sys.argv[1] if len(sys.argv) > 1 else "/root/reference"
    out_dir = sys.argv[2] if len(sys.argv) > 2 else os.path.join(os.path.dirname(__file__), "..", "tests", "golden")
    os.makedirs(out_dir, exist_ok=True)
    it = open(os.path.join(ref, "tests/integration_tests.rs")).read()
    step = it[it.index("fn test_aes_encryption_step_by_step"):it.index("fn test_encrypt_a_16_bytes_plaintext")]
    trace = {
        "source": "reference tests/integration_tests.rs:50-310 (FIPS-197 Appendix B)",
        "plaintext": arrays_after(step, "plaintext")[0],
        "key": arrays_after(step, "key")[0],
        "ciphertext": arrays_after(step, "expected_output")[0],
        "start_of_round": arrays_after(step, "expected_start_of_round"),
        "after_sub_bytes": arrays_after(step, "expected_after_substituting_bytes"),
        "after_shift_rows": arrays_after(step, "expected_after_shift_rows"),
        "after_mix_columns": arrays_after(step, "expected_after_mix_columns"),
    }
    # the table's 11th entry is the identifier `expected_output` (integration_tests.rs:118-119)
    trace["start_of_round"].append(trace["ciphertext"])
    assert len(trace["start_of_round"]) == 11 and len(trace["after_sub_bytes"]) == 10
    assert len(trace["after_shift_rows"]) == 10 and len(trace["after_mix_columns"]) == 10
    e2e = []
    for fn, nxt in (("fn test_encrypt_a_16_bytes_plaintext", "fn test_one_round_aes_encryption_of_a_64_bytes_plaintext"),
                    ("fn test_one_round_aes_encryption_of_a_64_bytes_plaintext", None)):
        seg = it[it.index(fn):it.index(nxt) if nxt else len(it)]
        e2e.append({
            "source": "reference tests/integration_tests.rs:%s" % ("313-337" if nxt else "340-372"),
            "plaintext": "".join(arrays_after(seg, "plaintext")),
            "key": arrays_after(seg, "key")[0],
            "ciphertext": "".join(arrays_after(seg, "expected_ciphertext")),
            "wrong_ciphertext": "".join(arrays_after(seg, "wrong_ciphertext")),
        })
    ac = open(os.path.join(ref, "src/aes_circuit.rs")).read()
    tests = ac[ac.index("mod tests"):]
    def seg(fn):
        i = tests.index("fn " + fn)
        j = tests.find("#[test]", i)
        return tests[i:j if j > 0 else len(tests)]
    def hexes(s):
        return [bytes(int(v, 16) for v in re.findall(r"0x([0-9a-fA-F]{1,2})", g)).hex()
                for g in re.findall(r"\[([^\[\]]*0x[^\[\]]*)\]", re.sub(r"//[^\n]*", "", s))]
    ark = hexes(seg("test_one_round_add_round_key_circuit"))
    mix = hexes(seg("test_one_round_column_mix_circuit"))
    sub = hexes(seg("test_one_round_sub_bytes_circuit"))
    kex = hexes(seg("key_expansion_circuit"))
    steps = {
        "source": "reference src/aes_circuit.rs:704-846",
        "add_round_key": {"input": ark[0], "key": ark[1], "output": ark[2]},
        "mix_columns": {"input": mix[0], "output": mix[1]},
        "sub_bytes": {"input": sub[0], "output": sub[1]},
        "key_expansion": {"key": kex[0], "round_key_10": kex[1]},
    }
    # S-box constants of lookup_table (src/aes_circuit.rs:433-694)
    lt = ac[ac.index("pub fn lookup_table"):ac.index("#[cfg(test)]")]
    sbox = [int(v, 16) for v in re.findall(r"new_constant\(cs(?:\.clone\(\))?,\s*0x([0-9a-fA-F]{2})\)", lt)]
    assert len(sbox) == 256
    steps["lookup_table"] = bytes(sbox).hex()
    # src/main.rs example
    from cryptography.hazmat.primitives.ciphers import Cipher, algorithms, modes
    enc = Cipher(algorithms.AES(bytes(16)), modes.ECB()).encryptor()
    main_ct = (enc.update(bytes([1] * 16)) + enc.finalize()).hex()
    e2e.append({"source": "reference src/main.rs:11-12 (ciphertext recomputed with python `cryptography`)",
                "plaintext": "01" * 16, "key": "00" * 16, "ciphertext": main_ct, "wrong_ciphertext": None})
    json.dump(trace, open(os.path.join(out_dir, "fips197_round_trace.json"), "w"), indent=1)
    json.dump(e2e, open(os.path.join(out_dir, "encrypt_e2e.json"), "w"), indent=1)
    json.dump(steps, open(os.path.join(out_dir, "gadget_steps.json"), "w"), indent=1)
    print("wrote 3 fixtures to", out_dir)

if __name__ == "__main__":
    main()
