// pipebench.cu -- issue-rate probe for the integer instructions the block encoders are built from.
// Diagnostic tool (not part of libgoofy_b200.so).  Every op is emitted through inline PTX and the
// eight chains feed each other, so neither nvvm nor ptxas can fold them.  Reports warp-instructions
// per cycle per SM sub-partition (SMSP); the SASS actually emitted is checked with cuobjdump.
//   build: make -C tools        run: tools/pipebench
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>

#define ITERS 1024

#define ROUND(OP) \
    OP(x0, x1, x2) OP(x1, x2, x3) OP(x2, x3, x4) OP(x3, x4, x5) OP(x4, x5, x6) OP(x5, x6, x7) OP(x6, x7, x0) OP(x7, x0, x1)
#define ROUND2(A, B) \
    A(x0, x1, x2) B(x1, x2, x3) A(x2, x3, x4) B(x3, x4, x5) A(x4, x5, x6) B(x5, x6, x7) A(x6, x7, x0) B(x7, x0, x1)

#define P_LOP3(d, a, b) asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(d) : "r"(a), "r"(b));
#define P_IADD3(d, a, b) asm volatile("{.reg .b32 t; add.u32 t, %0, %1; add.u32 %0, t, %2;}" : "+r"(d) : "r"(a), "r"(b));
#define P_IADD(d, a, b) asm volatile("add.u32 %0, %0, %1;" : "+r"(d) : "r"(a));
#define P_SHF(d, a, b) asm volatile("shf.l.wrap.b32 %0, %0, %1, 3;" : "+r"(d) : "r"(a));
#define P_SHR(d, a, b) asm volatile("shr.u32 %0, %0, 1; xor.b32 %0, %0, %1;" : "+r"(d) : "r"(a));
#define P_PRMT(d, a, b) asm volatile("prmt.b32 %0, %0, %1, 0x2503;" : "+r"(d) : "r"(a));
#define P_IMAD(d, a, b) asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(d) : "r"(a), "r"(b));
#define P_IDP(d, a, b) asm volatile("dp4a.u32.u32 %0, %0, %1, %2;" : "+r"(d) : "r"(a), "r"(b));
#define P_MNMX2(d, a, b) asm volatile("min.u16x2 %0, %0, %1;" : "+r"(d) : "r"(a));
#define P_MNMX3(d, a, b) asm volatile("{.reg .b32 t; min.u16x2 t, %0, %1; min.u16x2 %0, t, %2;}" : "+r"(d) : "r"(a), "r"(b));
#define P_ADDMNMX(d, a, b) asm volatile("{.reg .b32 t; add.s16x2 t, %0, %1; min.s16x2.relu %0, t, %2;}" : "+r"(d) : "r"(a), "r"(b));
#define P_VADD2(d, a, b) asm volatile("add.u16x2 %0, %0, %1;" : "+r"(d) : "r"(a));
#define P_LEAHI(d, a, b) asm volatile("{.reg .b32 t; shr.u32 t, %0, 1; add.u32 %0, t, %1;}" : "+r"(d) : "r"(a));
#define P_FADD(d, a, b) asm volatile("add.f32 %0, %0, %1;" : "+r"(d) : "r"(a));
#define P_FFMA(d, a, b) asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+r"(d) : "r"(a), "r"(b));
#define P_ABSD4(d, a, b) asm volatile("vabsdiff4.u32.u32.u32 %0, %0, %1, %2;" : "+r"(d) : "r"(a), "r"(b));
#define P_POPC(d, a, b) asm volatile("{.reg .b32 t; popc.b32 t, %0; xor.b32 %0, t, %1;}" : "+r"(d) : "r"(a));
#define P_IMADHI(d, a, b) asm volatile("mad.hi.u32 %0, %0, %1, %2;" : "+r"(d) : "r"(a), "r"(b));
#define P_IMADWIDE(d, a, b) asm volatile("{.reg .b64 t, u; mov.b64 u, {%0, %1}; mad.wide.u32 t, %0, %2, u; mov.b64 {%0, %1}, t;}" : "+r"(d), "+r"(a) : "r"(b));
#define P_LEA(d, a, b) asm volatile("{.reg .b32 t; shl.b32 t, %0, 3; add.u32 %0, t, %1;}" : "+r"(d) : "r"(a));
#define P_WIDE(d, a, b) asm volatile("{.reg .b64 t, u; mov.b64 u, {%0, %1}; mad.wide.u32 t, %2, %2, u; mov.b64 {%0, %1}, t;}" : "+r"(d), "+r"(a) : "r"(b));
#define P_HMNMX2(d, a, b) asm volatile("min.f16x2 %0, %0, %1;" : "+r"(d) : "r"(a));
#define P_FMNMX(d, a, b) asm volatile("min.f32 %0, %0, %1;" : "+r"(d) : "r"(a));
#define P_HFMA2(d, a, b) asm volatile("fma.rn.f16x2 %0, %0, %1, %2;" : "+r"(d) : "r"(a), "r"(b));
#define P_ISETP_SEL(d, a, b) asm volatile("{.reg .pred p; setp.lt.u32 p, %0, %1; selp.b32 %0, %1, %2, p;}" : "+r"(d) : "r"(a), "r"(b));

#define DEFINE_KERNEL(NAME, BODY)                                                                   \
    __global__ void __launch_bounds__(1024) NAME(uint32_t a, uint32_t b, uint32_t* out, unsigned long long* cyc) \
    {                                                                                               \
        uint32_t x0 = threadIdx.x, x1 = x0 * 3 + a, x2 = x0 * 5 + b, x3 = x0 * 7, x4 = x0 + 11,      \
                 x5 = x0 ^ a, x6 = x0 ^ b, x7 = x0 + a * b;                                          \
        __shared__ unsigned long long sStart, sEnd;                                                 \
        if (threadIdx.x == 0) { sStart = ~0ull; sEnd = 0ull; }                                      \
        __syncthreads();                                                                            \
        unsigned long long t0 = clock64();                                                          \
        _Pragma("unroll 4") for (int i = 0; i < ITERS; ++i) { BODY }                                  \
        unsigned long long t1 = clock64();                                                          \
        out[blockIdx.x * blockDim.x + threadIdx.x] = x0 ^ x1 ^ x2 ^ x3 ^ x4 ^ x5 ^ x6 ^ x7;           \
        /* whole-CTA span: first warp to start .. last warp to finish */                            \
        if ((threadIdx.x & 31) == 0) { atomicMin(&sStart, t0); atomicMax(&sEnd, t1); }              \
        __syncthreads();                                                                            \
        if (threadIdx.x == 0) cyc[blockIdx.x] = sEnd - sStart;                                      \
    }

DEFINE_KERNEL(k_lop3, ROUND(P_LOP3))
DEFINE_KERNEL(k_iadd, ROUND(P_IADD))
DEFINE_KERNEL(k_iadd3, ROUND(P_IADD3))
DEFINE_KERNEL(k_shf, ROUND(P_SHF))
DEFINE_KERNEL(k_prmt, ROUND(P_PRMT))
DEFINE_KERNEL(k_imad, ROUND(P_IMAD))
DEFINE_KERNEL(k_idp, ROUND(P_IDP))
DEFINE_KERNEL(k_mnmx2, ROUND(P_MNMX2))
DEFINE_KERNEL(k_mnmx3, ROUND(P_MNMX3))
DEFINE_KERNEL(k_addmnmx, ROUND(P_ADDMNMX))
DEFINE_KERNEL(k_vadd2, ROUND(P_VADD2))
DEFINE_KERNEL(k_leahi, ROUND(P_LEAHI))
DEFINE_KERNEL(k_fadd, ROUND(P_FADD))
DEFINE_KERNEL(k_ffma, ROUND(P_FFMA))
DEFINE_KERNEL(k_absd4, ROUND(P_ABSD4))
DEFINE_KERNEL(k_popc, ROUND(P_POPC))
DEFINE_KERNEL(k_setp_sel, ROUND(P_ISETP_SEL))
DEFINE_KERNEL(k_wide, ROUND(P_WIDE))
DEFINE_KERNEL(k_lop3_wide, ROUND2(P_LOP3, P_WIDE))
DEFINE_KERNEL(k_hmnmx2, ROUND(P_HMNMX2))
DEFINE_KERNEL(k_fmnmx, ROUND(P_FMNMX))
DEFINE_KERNEL(k_hfma2, ROUND(P_HFMA2))
DEFINE_KERNEL(k_lop3_hmnmx2, ROUND2(P_LOP3, P_HMNMX2))
DEFINE_KERNEL(k_imad_hmnmx2, ROUND2(P_IMAD, P_HMNMX2))
DEFINE_KERNEL(k_lop3_fmnmx, ROUND2(P_LOP3, P_FMNMX))
DEFINE_KERNEL(k_imadhi, ROUND(P_IMADHI))
DEFINE_KERNEL(k_lea, ROUND(P_LEA))
DEFINE_KERNEL(k_lop3_imadhi, ROUND2(P_LOP3, P_IMADHI))
DEFINE_KERNEL(k_lop3_lea, ROUND2(P_LOP3, P_LEA))
DEFINE_KERNEL(k_lop3_imad, ROUND2(P_LOP3, P_IMAD))
DEFINE_KERNEL(k_lop3_idp, ROUND2(P_LOP3, P_IDP))
DEFINE_KERNEL(k_imad_idp, ROUND2(P_IMAD, P_IDP))
DEFINE_KERNEL(k_lop3_mnmx2, ROUND2(P_LOP3, P_MNMX2))
DEFINE_KERNEL(k_lop3_mnmx3, ROUND2(P_LOP3, P_MNMX3))
DEFINE_KERNEL(k_imad_mnmx2, ROUND2(P_IMAD, P_MNMX2))
DEFINE_KERNEL(k_imad_mnmx3, ROUND2(P_IMAD, P_MNMX3))
DEFINE_KERNEL(k_lop3_fadd, ROUND2(P_LOP3, P_FADD))
DEFINE_KERNEL(k_imad_fadd, ROUND2(P_IMAD, P_FADD))
DEFINE_KERNEL(k_idp_ffma, ROUND2(P_IDP, P_FFMA))
DEFINE_KERNEL(k_lop3_prmt, ROUND2(P_LOP3, P_PRMT))
DEFINE_KERNEL(k_lop3_shf, ROUND2(P_LOP3, P_SHF))
DEFINE_KERNEL(k_lop3_vadd2, ROUND2(P_LOP3, P_VADD2))
DEFINE_KERNEL(k_imad_vadd2, ROUND2(P_IMAD, P_VADD2))
DEFINE_KERNEL(k_lop3_iadd, ROUND2(P_LOP3, P_IADD))
DEFINE_KERNEL(k_imad_iadd, ROUND2(P_IMAD, P_IADD))
DEFINE_KERNEL(k_lop3_addmnmx, ROUND2(P_LOP3, P_ADDMNMX))
DEFINE_KERNEL(k_imad_addmnmx, ROUND2(P_IMAD, P_ADDMNMX))

typedef void (*kern_t)(uint32_t, uint32_t, uint32_t*, unsigned long long*);

static void run(const char* name, kern_t k, int warpsPerSmsp, int sms, double instrPerOp)
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
    double ops = (double)ITERS * 8.0 * warpsPerSmsp;  // source-level ops per SMSP
    printf("%-14s warps/smsp=%d  cycles=%9.0f  ops/clk/SMSP=%.3f\n", name, warpsPerSmsp, avg, ops / avg);
    cudaFree(out); cudaFree(cyc);
}

int main()
{
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    printf("device %s, %d SMs, L2 %d MB  (ops = source-level ops; see SASS for how each lowers)\n", p.name,
           p.multiProcessorCount, p.l2CacheSize >> 20);
    int sms = p.multiProcessorCount;
    struct { const char* n; kern_t k; } ks[] = {
        {"lop3", k_lop3}, {"iadd", k_iadd}, {"iadd3", k_iadd3}, {"shf", k_shf}, {"prmt", k_prmt}, {"imad", k_imad},
        {"idp4a", k_idp}, {"vimnmx.u16x2", k_mnmx2}, {"vimnmx3.u16x2", k_mnmx3}, {"viaddmnmx", k_addmnmx},
        {"viadd.16x2", k_vadd2}, {"(x>>1)+a", k_leahi}, {"fadd", k_fadd}, {"ffma", k_ffma}, {"vabsdiff4", k_absd4},
        {"popc+xor", k_popc}, {"setp+selp", k_setp_sel},
        {"imad.wide", k_wide}, {"lop3+imad.wide", k_lop3_wide}, {"hmnmx2", k_hmnmx2}, {"fmnmx", k_fmnmx}, {"hfma2", k_hfma2},
        {"lop3+hmnmx2", k_lop3_hmnmx2}, {"imad+hmnmx2", k_imad_hmnmx2}, {"lop3+fmnmx", k_lop3_fmnmx},
        {"imad.hi", k_imadhi}, {"(x<<3)+a", k_lea}, {"lop3+imad.hi", k_lop3_imadhi}, {"lop3+(x<<3)+a", k_lop3_lea},
        {"lop3+imad", k_lop3_imad}, {"lop3+idp", k_lop3_idp}, {"imad+idp", k_imad_idp},
        {"lop3+mnmx2", k_lop3_mnmx2}, {"lop3+mnmx3", k_lop3_mnmx3}, {"imad+mnmx2", k_imad_mnmx2}, {"imad+mnmx3", k_imad_mnmx3},
        {"lop3+fadd", k_lop3_fadd}, {"imad+fadd", k_imad_fadd}, {"idp+ffma", k_idp_ffma}, {"lop3+prmt", k_lop3_prmt},
        {"lop3+shf", k_lop3_shf}, {"lop3+viadd16x2", k_lop3_vadd2}, {"imad+viadd16x2", k_imad_vadd2},
        {"lop3+iadd", k_lop3_iadd}, {"imad+iadd", k_imad_iadd}, {"lop3+addmnmx", k_lop3_addmnmx}, {"imad+addmnmx", k_imad_addmnmx},
    };
    for (auto& e : ks) run(e.n, e.k, 8, sms, 1.0);
    return 0;
}
