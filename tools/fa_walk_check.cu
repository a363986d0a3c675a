// Host-side check of the bit-parallel FASTA walk (kb_fa_walk_bits, csrc/kb_ingest.cuh) against its byte-at-a-time form
// (kb_fa_walk_bytes) on random 16-byte chunks: same keep / separator masks, flags, end state.  No GPU needed:
//   nvcc -std=c++17 -I krisp_b200/csrc -o /tmp/fa_walk_check tools/fa_walk_check.cu && /tmp/fa_walk_check
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include "kb_ingest.cuh"

int main() {
    const char alpha[] = "ACGTNacgtn\n\n\n\r>>  \tUu\x0b\x0c-*RY\x01\x7f\x80\xff";
    const int na = (int)sizeof(alpha) - 1;
    uint64_t rng = 88172645463325252ULL;
    auto next = [&]() { rng ^= rng << 13; rng ^= rng >> 7; rng ^= rng << 17; return rng; };
    long bad = 0, n = 0;
    for (long it = 0; it < 4000000; it++) {
        uint8_t b[KB_FA_PER];
        const int mode = (int)(next() % 4);                       // 0: mostly bases, 1: dense specials, 2: header-like, 3: CRLF lines
        for (int i = 0; i < KB_FA_PER; i++) {
            const uint64_t r = next();
            if (mode == 0) b[i] = (r % 20) ? (uint8_t)"ACGT"[r % 4] : (uint8_t)alpha[(r >> 8) % na];
            else if (mode == 1) b[i] = (uint8_t)alpha[r % na];
            else if (mode == 2) b[i] = (r % 6 == 0) ? (uint8_t)'\n' : ((r % 6 == 1) ? (uint8_t)'>' : (uint8_t)alpha[(r >> 8) % na]);
            else b[i] = (r % 8 == 0) ? (uint8_t)'\r' : ((r % 8 == 1) ? (uint8_t)'\n' : (uint8_t)"ACGTUu> "[(r >> 8) % 8]);
        }
        const uint32_t nv = (next() % 5) ? 16u : (uint32_t)(next() % 17);
        const uint8_t prevs[] = {'\n', 'A', '>', '\r'}, nexts[] = {'\n', 'A', '\r', '>'};
        const uint8_t prev = prevs[next() % 4], nx = nexts[next() % 4];
        const int fasta = (int)(next() & 1);
        const uint32_t state = fasta ? (uint32_t)(next() & 1) : 0u;
        uint32_t w[4];
        for (int j = 0; j < 4; j++) {
            w[j] = 0;
            for (int i = 0; i < 4; i++) w[j] |= (uint32_t)((4 * j + i) < (int)nv ? b[4 * j + i] : (uint8_t)'\n') << (8 * i);
        }
        const KbFaWalk ref = kb_fa_walk_bytes(fasta, b, nv, prev, nx, state);
        const KbFaBits m = kb_fa_bits(w, nv, prev == '\n', fasta);
        const KbFaWalk got = kb_fa_walk_bits(m, w, nv, nx == '\n', fasta, state);
        n++;
        if (ref.keep != got.keep || ref.sep != got.sep || ref.flags != got.flags || ref.end_state != got.end_state || ref.has_ls != got.has_ls) {
            if (bad++ < 10) {
                printf("MISMATCH nv=%u prev=%02x next=%02x fasta=%d state=%u bytes=", nv, prev, nx, fasta, state);
                for (int i = 0; i < 16; i++) printf("%02x ", b[i]);
                printf("\n  ref keep=%04x sep=%04x fl=%u end=%u ls=%u\n  got keep=%04x sep=%04x fl=%u end=%u ls=%u\n",
                       ref.keep, ref.sep, ref.flags, ref.end_state, ref.has_ls, got.keep, got.sep, got.flags, got.end_state, got.has_ls);
            }
        }
    }
    printf("%ld chunks, %ld mismatches\n", n, bad);
    return bad ? 1 : 0;
}
