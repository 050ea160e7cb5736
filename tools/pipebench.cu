// pipebench.cu -- issue-rate probe for the integer instructions the block encoders are built from.
// Diagnostic tool (not part of libgoofy_b200.so).  For each op it runs 8 independent dependency
// chains per thread and reports warp-instructions per cycle per SM sub-partition, so the
// instruction budget in DESIGN.md can be stated against measured B200 pipe rates.
//   build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/pipebench tools/pipebench.cu
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>
#include <cstring>

#define ITERS 2048

struct Result { unsigned long long cycles; };

#define CHAINS8(EXPR)                                                         \
    x0 = EXPR(x0); x1 = EXPR(x1); x2 = EXPR(x2); x3 = EXPR(x3);               \
    x4 = EXPR(x4); x5 = EXPR(x5); x6 = EXPR(x6); x7 = EXPR(x7);

#define DEFINE_KERNEL(NAME, BODY)                                                                   \
    __global__ void __launch_bounds__(1024) NAME(uint32_t a, uint32_t b, uint32_t* out, unsigned long long* cyc) \
    {                                                                                               \
        uint32_t x0 = threadIdx.x, x1 = x0 * 3 + a, x2 = x0 * 5 + b, x3 = x0 * 7, x4 = x0 + 11,      \
                 x5 = x0 ^ a, x6 = x0 ^ b, x7 = x0 + a * b;                                          \
        unsigned long long t0 = clock64();                                                          \
        _Pragma("unroll 4") for (int i = 0; i < ITERS; ++i) { BODY }                                  \
        unsigned long long t1 = clock64();                                                          \
        out[blockIdx.x * blockDim.x + threadIdx.x] = x0 ^ x1 ^ x2 ^ x3 ^ x4 ^ x5 ^ x6 ^ x7;           \
        if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;                                            \
    }

#define OP_LOP3(x) (((x) & a) ^ b)
#define OP_IADD(x) ((x) + a)
#define OP_IADD3(x) ((x) + a + b)
#define OP_SHF(x) __funnelshift_l((x), b, 3)
#define OP_PRMT(x) __byte_perm((x), a, 0x2103)
#define OP_IMAD(x) ((x) * a + b)
#define OP_IDP(x) __dp4a((x), a, b)
#define OP_MNMX3(x) __vimin3_u16x2((x), a, b)
#define OP_MNMX2(x) __vminu2((x), a)
#define OP_ADDMNMX(x) __viaddmin_s16x2_relu((x), a, b)
#define OP_LEAHI(x) (((x) >> 1) + a)
#define OP_SHL8(x) ((x) << 8)
#define OP_FADD(x) __float_as_uint(__uint_as_float(x) + __uint_as_float(a))
#define OP_ABSDIFF4(x) __vabsdiffu4((x), a)
#define OP_VADD2(x) __vadd2((x), a)

DEFINE_KERNEL(k_lop3, CHAINS8(OP_LOP3))
DEFINE_KERNEL(k_iadd, CHAINS8(OP_IADD))
DEFINE_KERNEL(k_iadd3, CHAINS8(OP_IADD3))
DEFINE_KERNEL(k_shf, CHAINS8(OP_SHF))
DEFINE_KERNEL(k_prmt, CHAINS8(OP_PRMT))
DEFINE_KERNEL(k_imad, CHAINS8(OP_IMAD))
DEFINE_KERNEL(k_idp, CHAINS8(OP_IDP))
DEFINE_KERNEL(k_mnmx3, CHAINS8(OP_MNMX3))
DEFINE_KERNEL(k_mnmx2, CHAINS8(OP_MNMX2))
DEFINE_KERNEL(k_addmnmx, CHAINS8(OP_ADDMNMX))
DEFINE_KERNEL(k_leahi, CHAINS8(OP_LEAHI))
DEFINE_KERNEL(k_fadd, CHAINS8(OP_FADD))
DEFINE_KERNEL(k_absdiff4, CHAINS8(OP_ABSDIFF4))
DEFINE_KERNEL(k_vadd2, CHAINS8(OP_VADD2))
// mixes: 4 chains of one op + 4 of another
#define MIX(A, B) x0 = A(x0); x1 = B(x1); x2 = A(x2); x3 = B(x3); x4 = A(x4); x5 = B(x5); x6 = A(x6); x7 = B(x7);
DEFINE_KERNEL(k_lop3_imad, MIX(OP_LOP3, OP_IMAD))
DEFINE_KERNEL(k_lop3_idp, MIX(OP_LOP3, OP_IDP))
DEFINE_KERNEL(k_imad_idp, MIX(OP_IMAD, OP_IDP))
DEFINE_KERNEL(k_lop3_mnmx3, MIX(OP_LOP3, OP_MNMX3))
DEFINE_KERNEL(k_imad_mnmx3, MIX(OP_IMAD, OP_MNMX3))
DEFINE_KERNEL(k_lop3_fadd, MIX(OP_LOP3, OP_FADD))
DEFINE_KERNEL(k_imad_fadd, MIX(OP_IMAD, OP_FADD))
DEFINE_KERNEL(k_lop3_prmt, MIX(OP_LOP3, OP_PRMT))
DEFINE_KERNEL(k_lop3_shf, MIX(OP_LOP3, OP_SHF))
DEFINE_KERNEL(k_idp_mnmx3, MIX(OP_IDP, OP_MNMX3))

typedef void (*kern_t)(uint32_t, uint32_t, uint32_t*, unsigned long long*);

static void run(const char* name, kern_t k, int warpsPerSmsp, int sms)
{
    uint32_t* out; unsigned long long* cyc;
    int threads = warpsPerSmsp * 4 * 32;
    cudaMalloc(&out, (size_t)sms * threads * 4);
    cudaMalloc(&cyc, sms * 8);
    k<<<sms, threads>>>(0x01020304u, 0x00FF00FFu, out, cyc);
    k<<<sms, threads>>>(0x01020304u, 0x00FF00FFu, out, cyc);
    cudaDeviceSynchronize();
    unsigned long long h[1024];
    cudaMemcpy(h, cyc, sms * 8, cudaMemcpyDeviceToHost);
    double avg = 0; for (int i = 0; i < sms; ++i) avg += (double)h[i]; avg /= sms;
    double instr = (double)ITERS * 8.0 * warpsPerSmsp;  // warp-instructions per SMSP
    printf("%-14s warps/smsp=%d  cycles=%9.0f  warp-instr/clk/SMSP=%.3f\n", name, warpsPerSmsp, avg, instr / avg);
    cudaFree(out); cudaFree(cyc);
}

int main()
{
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    printf("device %s, %d SMs, L2 %d MB\n", p.name, p.multiProcessorCount, p.l2CacheSize >> 20);
    int sms = p.multiProcessorCount;
    struct { const char* n; kern_t k; } ks[] = {
        {"lop3", k_lop3}, {"iadd", k_iadd}, {"iadd3", k_iadd3}, {"shf", k_shf}, {"prmt", k_prmt}, {"imad", k_imad},
        {"idp4a", k_idp}, {"vimnmx3.u16x2", k_mnmx3}, {"vimnmx.u16x2", k_mnmx2}, {"viaddmnmx", k_addmnmx},
        {"(x>>1)+a", k_leahi}, {"fadd", k_fadd}, {"vabsdiff4", k_absdiff4}, {"vadd2", k_vadd2},
        {"lop3+imad", k_lop3_imad}, {"lop3+idp", k_lop3_idp}, {"imad+idp", k_imad_idp}, {"lop3+mnmx3", k_lop3_mnmx3},
        {"imad+mnmx3", k_imad_mnmx3}, {"lop3+fadd", k_lop3_fadd}, {"imad+fadd", k_imad_fadd}, {"lop3+prmt", k_lop3_prmt},
        {"lop3+shf", k_lop3_shf}, {"idp+mnmx3", k_idp_mnmx3},
    };
    for (auto& e : ks) { run(e.n, e.k, 4, sms); }
    for (auto& e : ks) { run(e.n, e.k, 8, sms); }
    return 0;
}
