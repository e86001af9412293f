// Host build of the generated limb arithmetic (csrc/fp_gen.inc): the C emulation of the exact
// PTX instruction list the GPU runs.  Test-only.
#include <cstdio>
#include <cstdlib>
#define ZK_EMU_ASSERT(x) do { if (!(x)) { fprintf(stderr, "emu assert failed: %s line %d\n", #x, __LINE__); abort(); } } while (0)
#include "fp_gen.inc"
extern "C" {
// op 0 mul 1 add 2 sub 3 sqr; field 0 fr 1 fq; n elements of 8 u32
void emu_field_op(int field, int op, const uint32_t* a, const uint32_t* b, uint32_t* out, size_t n) {
    for (size_t i = 0; i < n; ++i) {
        const uint32_t* x = a + 8 * i; const uint32_t* y = b ? b + 8 * i : nullptr; uint32_t* r = out + 8 * i;
        if (field == 0) { switch (op) { case 0: fr_mul(r, x, y); break; case 1: fr_add(r, x, y); break; case 2: fr_sub(r, x, y); break; default: fr_sqr(r, x); } }
        else            { switch (op) { case 0: fq_mul(r, x, y); break; case 1: fq_add(r, x, y); break; case 2: fq_sub(r, x, y); break; default: fq_sqr(r, x); } }
    }
}
// out = a*b + c*d with one reduction (fe_mul_add2); carry-is-zero assertions of the dual product run inside
void emu_mul_add2(int field, const uint32_t* a, const uint32_t* b, const uint32_t* c, const uint32_t* d, uint32_t* out, size_t n) {
    for (size_t i = 0; i < n; ++i) {
        if (field == 0) fr_mul_add2(out + 8 * i, a + 8 * i, b + 8 * i, c + 8 * i, d + 8 * i);
        else fq_mul_add2(out + 8 * i, a + 8 * i, b + 8 * i, c + 8 * i, d + 8 * i);
    }
}
}
