/* C (not C++) consumer of include/zkaes_b200.h: what a cgo / Rust -sys / JNI binding sees.  Built by tests/test_abi.py with
 * gcc -std=c99 against libzkaes_b200.so and run on the golden proof: argv = key file, proof file, ciphertext file.
 * Exit code 0 = accepted, 1 = rejected, 2 = error.  Test infrastructure only. */
#include <stdio.h>
#include <stdlib.h>

#include "../include/zkaes_b200.h"

static unsigned char* slurp(const char* path, size_t* len) {
    FILE* f = fopen(path, "rb");
    if (!f) return NULL;
    fseek(f, 0, SEEK_END);
    long n = ftell(f);
    fseek(f, 0, SEEK_SET);
    unsigned char* buf = (unsigned char*)malloc(n > 0 ? (size_t)n : 1);
    if (fread(buf, 1, (size_t)n, f) != (size_t)n) {
        fclose(f);
        free(buf);
        return NULL;
    }
    fclose(f);
    *len = (size_t)n;
    return buf;
}

int main(int argc, char** argv) {
    if (argc != 4) return 2;
    size_t vk_len = 0, proof_len = 0, ct_len = 0;
    unsigned char* vk = slurp(argv[1], &vk_len);
    unsigned char* proof = slurp(argv[2], &proof_len);
    unsigned char* ct = slurp(argv[3], &ct_len);
    if (!vk || !proof || !ct) return 2;
    int accepted = -1;
    int rc = zkaes_verify_encryption(vk, vk_len, proof, proof_len, ct, ct_len, &accepted);
    if (rc != 0) {
        fprintf(stderr, "error %d: %s\n", rc, zkaes_last_error(NULL));
        return 2;
    }
    printf("%s\n", accepted ? "accepted" : "rejected");
    return accepted ? 0 : 1;
}
