// tools/int32_pipes.cu -- issue-rate micro-benchmarks of the integer instructions the NW
// kernels are built from, on every SM.  Prints lane-ops per clock per SM for each.
//   nvcc --cudart=shared -gencode arch=compute_100a,code=sm_100a -O3 -o int32_pipes int32_pipes.cu && ./int32_pipes
// Each kernel runs 16 independent dependency chains per thread; operands live in registers
// that ptxas cannot prove equal, so chains are not merged or hoisted (checked in the SASS).
#include <cstdio>
#include <cuda_runtime.h>

#define CHAINS 16
#define BODY(NAME, STMT)                                                                   \
    __global__ void __launch_bounds__(256) NAME(int iters, const int *__restrict__ src,    \
                                                int *sink)                                 \
    {                                                                                      \
        int x[CHAINS], a[CHAINS], b[CHAINS];                                               \
        _Pragma("unroll") for (int j = 0; j < CHAINS; ++j) {                               \
            x[j] = src[threadIdx.x + j * 7];                                               \
            a[j] = src[threadIdx.x + j * 5 + 1];                                           \
            b[j] = src[threadIdx.x + j * 3 + 2];                                           \
        }                                                                                  \
        for (int it = 0; it < iters; ++it) {                                               \
            _Pragma("unroll") for (int j = 0; j < CHAINS; ++j) { STMT; }                   \
        }                                                                                  \
        int acc = 0;                                                                       \
        _Pragma("unroll") for (int j = 0; j < CHAINS; ++j) acc ^= x[j] ^ a[j] ^ b[j];      \
        if (acc == 0x7fffffff) sink[0] = acc;                                              \
    }

// two instructions per chain per iteration unless noted
BODY(k_iadd,     asm volatile("add.s32 %0, %0, %1;" : "+r"(x[j]) : "r"(a[j]));
                 asm volatile("add.s32 %0, %0, %1;" : "+r"(x[j]) : "r"(b[j])))
BODY(k_minmax,   asm volatile("max.s32 %0, %0, %1;" : "+r"(x[j]) : "r"(a[j]));
                 asm volatile("min.s32 %0, %0, %1;" : "+r"(x[j]) : "r"(b[j])))
BODY(k_addmax_sep, asm volatile("add.s32 %0, %0, %1;" : "+r"(x[j]) : "r"(a[j]));
                 asm volatile("min.s32 %0, %0, %1;" : "+r"(x[j]) : "r"(b[j])))
BODY(k_viaddmnmx, x[j] = __viaddmax_s32(x[j], a[j], b[j]); x[j] = __viaddmin_s32(x[j], b[j], a[j]))
BODY(k_vimnmx3,  x[j] = __vimax3_s32(x[j], a[j], b[j]); x[j] = __vimin3_s32(x[j], b[j], a[j] ^ 5))
BODY(k_lop3,     asm volatile("lop3.b32 %0, %0, %1, %2, 0xE8;" : "+r"(x[j]) : "r"(a[j]), "r"(b[j]));
                 asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(x[j]) : "r"(b[j]), "r"(a[j])))
BODY(k_imad,     asm volatile("mad.lo.s32 %0, %0, %1, %2;" : "+r"(x[j]) : "r"(a[j]), "r"(b[j]));
                 asm volatile("mad.lo.s32 %0, %0, %1, %2;" : "+r"(x[j]) : "r"(b[j]), "r"(a[j])))
BODY(k_prmt,     asm volatile("prmt.b32 %0, %0, %1, %2;" : "+r"(x[j]) : "r"(a[j]), "r"(b[j]));
                 asm volatile("prmt.b32 %0, %0, %1, %2;" : "+r"(x[j]) : "r"(b[j]), "r"(a[j])))
BODY(k_setp_sel, { int t; asm volatile("{ .reg .pred p; setp.gt.s32 p, %1, %2; selp.s32 %0, %3, %1, p; }"
                                       : "=r"(t) : "r"(x[j]), "r"(a[j]), "r"(b[j])); x[j] = t; })
BODY(k_shl_imad, asm volatile("shl.b32 %0, %0, 2;" : "+r"(x[j]));
                 asm volatile("xor.b32 %0, %0, %1;" : "+r"(x[j]) : "r"(a[j])))
BODY(k_mix_alu_fma, asm volatile("max.s32 %0, %0, %1;" : "+r"(x[j]) : "r"(a[j]));
                 asm volatile("mad.lo.s32 %0, %0, %1, %2;" : "+r"(x[j]) : "r"(b[j]), "r"(a[j])))
BODY(k_vimax_s16x2, x[j] = __vmaxs2(x[j], a[j]); x[j] = __vmins2(x[j], b[j]))
BODY(k_viaddmax_s16x2, x[j] = __viaddmax_s16x2(x[j], a[j], b[j]); x[j] = __viaddmin_s16x2(x[j], b[j], a[j]))
BODY(k_vadd2,    x[j] = __vadd2(x[j], a[j]); x[j] = __vsub2(x[j], b[j]))
BODY(k_shfl,     x[j] = __shfl_up_sync(0xffffffffu, x[j], 1); x[j] ^= a[j])

// add of a kernel-parameter constant (ptxas emits VIADD R, R, UR) -- which pipe is VIADD on?
__global__ void __launch_bounds__(256) k_viadd(int iters, const int *__restrict__ src, int *sink, int c1, int c2)
{
    int x[CHAINS];
    _Pragma("unroll") for (int j = 0; j < CHAINS; ++j) x[j] = src[threadIdx.x + j * 7];
    for (int it = 0; it < iters; ++it) {
        _Pragma("unroll") for (int j = 0; j < CHAINS; ++j) {
            asm volatile("add.s32 %0, %0, %1;" : "+r"(x[j]) : "r"(c1));
            asm volatile("xor.b32 %0, %0, %1;" : "+r"(x[j]) : "r"(c2));
        }
    }
    int acc = 0;
    _Pragma("unroll") for (int j = 0; j < CHAINS; ++j) acc ^= x[j];
    if (acc == 0x7fffffff) sink[0] = acc;
}
__global__ void __launch_bounds__(256) k_viadd_imad(int iters, const int *__restrict__ src, int *sink, int c1, int c2)
{
    int x[CHAINS], a[CHAINS];
    _Pragma("unroll") for (int j = 0; j < CHAINS; ++j) { x[j] = src[threadIdx.x + j * 7]; a[j] = src[threadIdx.x + j * 5 + 1]; }
    for (int it = 0; it < iters; ++it) {
        _Pragma("unroll") for (int j = 0; j < CHAINS; ++j) {
            asm volatile("add.s32 %0, %0, %1;" : "+r"(x[j]) : "r"(c1));
            asm volatile("mad.lo.s32 %0, %0, %1, %2;" : "+r"(x[j]) : "r"(a[j]), "r"(c2));
        }
    }
    int acc = 0;
    _Pragma("unroll") for (int j = 0; j < CHAINS; ++j) acc ^= x[j];
    if (acc == 0x7fffffff) sink[0] = acc;
}
__global__ void __launch_bounds__(256) k_lds_lop(int iters, const int *__restrict__ src, int *sink, int c1, int c2)
{
    __shared__ int tab[32 * 64];
    for (int i = threadIdx.x; i < 32 * 64; i += blockDim.x) tab[i] = src[i & 1023];
    __syncthreads();
    int x[CHAINS];
    _Pragma("unroll") for (int j = 0; j < CHAINS; ++j) x[j] = src[threadIdx.x + j * 7] & 63;
    const int lane = threadIdx.x & 31;
    for (int it = 0; it < iters; ++it) {
        _Pragma("unroll") for (int j = 0; j < CHAINS; ++j) {
            x[j] = tab[(x[j] & 63) * 32 + lane] ^ c2;       // conflict-free LDS + LOP3
        }
    }
    int acc = 0;
    _Pragma("unroll") for (int j = 0; j < CHAINS; ++j) acc ^= x[j];
    if (acc == 0x7fffffff) sink[0] = acc;
}
typedef void (*kern5_t)(int, const int *, int *, int, int);

typedef void (*kern_t)(int, const int *, int *);

int main()
{
    cudaDeviceProp prop;
    cudaGetDeviceProperties(&prop, 0);
    int clk_khz = 0;
    cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, 0);
    printf("# %s, %d SMs, max clock %d MHz\n", prop.name, prop.multiProcessorCount, clk_khz / 1000);
    int *src, *sink;
    cudaMalloc(&src, 4096 * sizeof(int));
    cudaMalloc(&sink, 64);
    cudaMemset(src, 1, 4096 * sizeof(int));
    struct { const char *name; kern_t fn; double instr_per_chain; } tests[] = {
        {"IADD3 (2 adds merged)", k_iadd, 1}, {"VIMNMX (max,min)", k_minmax, 2},
        {"add+min -> VIADDMNMX", k_addmax_sep, 1}, {"VIADDMNMX (x2)", k_viaddmnmx, 2},
        {"VIMNMX3 (x2)", k_vimnmx3, 2}, {"LOP3 (x2)", k_lop3, 2}, {"IMAD (x2)", k_imad, 2},
        {"PRMT (x2)", k_prmt, 2}, {"ISETP+SEL", k_setp_sel, 2}, {"SHL + XOR", k_shl_imad, 2},
        {"VIMNMX + IMAD", k_mix_alu_fma, 2}, {"VIMNMX.S16x2 (x2)", k_vimax_s16x2, 2},
        {"VIADDMNMX.S16x2 (x2)", k_viaddmax_s16x2, 2}, {"VADD2/VSUB2", k_vadd2, 2},
        {"SHFL.UP + XOR", k_shfl, 2},
    };
    const int iters = 1 << 13, threads = 256, blocks = prop.multiProcessorCount * 8;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    printf("%-24s %12s %14s %16s\n", "test", "ms", "Tinstr-lane/s", "lanes/clk/SM@max");
    for (auto &t : tests) {
        float best = 1e30f;
        for (int rep = 0; rep < 4; ++rep) {
            cudaEventRecord(e0);
            t.fn<<<blocks, threads>>>(iters, src, sink);
            cudaEventRecord(e1);
            cudaEventSynchronize(e1);
            float ms;
            cudaEventElapsedTime(&ms, e0, e1);
            if (rep > 0 && ms < best) best = ms;
        }
        double lanes = (double)blocks * threads * iters * CHAINS * t.instr_per_chain;
        double rate = lanes / (best * 1e-3);
        printf("%-24s %12.3f %14.2f %16.1f\n", t.name, best, rate / 1e12,
               rate / (prop.multiProcessorCount * (clk_khz * 1e3)));
    }
    struct { const char *name; kern5_t fn; double instr_per_chain; } tests5[] = {
        {"VIADD(UR) + LOP3", k_viadd, 2}, {"VIADD(UR) + IMAD", k_viadd_imad, 2}, {"LDS + 2xLOP3 + IMAD?", k_lds_lop, 1},
    };
    for (auto &t : tests5) {
        float best = 1e30f;
        for (int rep = 0; rep < 4; ++rep) {
            cudaEventRecord(e0);
            t.fn<<<blocks, threads>>>(iters, src, sink, 12345, 777);
            cudaEventRecord(e1);
            cudaEventSynchronize(e1);
            float ms;
            cudaEventElapsedTime(&ms, e0, e1);
            if (rep > 0 && ms < best) best = ms;
        }
        double lanes = (double)blocks * threads * iters * CHAINS * t.instr_per_chain;
        double rate = lanes / (best * 1e-3);
        printf("%-24s %12.3f %14.2f %16.1f\n", t.name, best, rate / 1e12,
               rate / (prop.multiProcessorCount * (clk_khz * 1e3)));
    }
    cudaError_t e = cudaDeviceSynchronize();
    printf("# status: %s\n", cudaGetErrorString(e));
    return e != cudaSuccess;
}
