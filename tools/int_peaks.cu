// int_peaks.cu — on-box microbenchmark of the sm_100a integer issue rates the SAD
// disparity kernels are bounded by ("P_int must be confirmed by an
// on-box microbenchmark"; MEASURED_PEAKS.json carries only HBM and bf16 peaks).
//
// Each test runs ITERS iterations of an unrolled body of CHAINS independent
// dependency chains of one instruction (or a fixed mix), on 148*k CTAs x 1024 threads,
// and reports lane-ops per clock per SM from clock64() deltas (frequency independent)
// and lane-ops per second from CUDA events.
//
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -o int_peaks int_peaks.cu
// Output: one JSON object on stdout.
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { \
  fprintf(stderr, "CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1);} } while (0)

constexpr int ITERS = 512;
constexpr int CHAINS = 8;

enum Op { OP_IADD3, OP_VABSDIFF4, OP_VABSDIFF4_ACC, OP_PRMT, OP_LOP3, OP_IMAD, OP_VIADD16X2,
          OP_VIMNMX16X2, OP_VIMNMX, OP_VIMNMX3, OP_DP4A, OP_SHF, OP_MIX_IADD3_IMAD,
          OP_MIX_VABS_IMAD, OP_MIX_PRMT_IMAD, OP_MIX_VIMNMX_IMAD, OP_LDS128, OP_MIX_LDS128_IADD3,
          OP_MIX_ALU3_IMAD1, OP_REDUX_MIN, OP_SHFL_BFLY, OP_MIX_REDUX_8ALU, OP_MIX_SHFL_8ALU, OP_COUNT };

static const char* op_name[OP_COUNT] = {
  "iadd3", "vabsdiff4", "vabsdiff4_acc", "prmt", "lop3", "imad", "viadd_16x2",
  "vimnmx_u16x2+lop3", "vimnmx_u32(min,max)", "vimnmx3_u32+viadd", "idp4a", "shf_funnel", "mix_iadd3+imad",
  "mix_vabsdiff4+imad", "mix_prmt+imad", "mix_vimnmx+imad", "lds128+lop3", "mix_lds128+4iadd3",
  "mix_3alu+1imad", "redux_min_u32", "shfl_bfly", "mix_redux+8iadd3", "mix_shfl+8iadd3" };
// instructions issued per chain per iteration
static const int op_instr[OP_COUNT] = {1,1,1,1,1,1,1,2,2,2,1,1,2,2,2,2,2,5,4,1,1,9,9};

template <int OP>
__device__ __forceinline__ void step(uint32_t& a, uint32_t b, uint32_t c, const uint4* sm, uint32_t& addr) {
  if constexpr (OP == OP_IADD3)        asm volatile("add.u32 %0, %0, %1;\n\tadd.u32 %0, %0, %2;" : "+r"(a) : "r"(b), "r"(c));
  if constexpr (OP == OP_VABSDIFF4)    asm volatile("vabsdiff4.u32.u32.u32 %0, %0, %1, %2;" : "+r"(a) : "r"(b), "r"(c));
  if constexpr (OP == OP_VABSDIFF4_ACC)asm volatile("vabsdiff4.u32.u32.u32.add %0, %1, %2, %0;" : "+r"(a) : "r"(b), "r"(c));
  if constexpr (OP == OP_PRMT)         asm volatile("prmt.b32 %0, %0, %1, 0x5140;" : "+r"(a) : "r"(b));
  if constexpr (OP == OP_LOP3)         asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(a) : "r"(b), "r"(c));
  if constexpr (OP == OP_IMAD)         asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(a) : "r"(b), "r"(c));
  if constexpr (OP == OP_VIADD16X2)    a = __vadd2(a, b);
  if constexpr (OP == OP_VIMNMX16X2)   a = __vminu2(a, b) ^ c;
  if constexpr (OP == OP_VIMNMX)       asm volatile("min.u32 %0, %0, %1;\n\tmax.u32 %0, %0, %2;" : "+r"(a) : "r"(b), "r"(c));
  if constexpr (OP == OP_VIMNMX3)      a = min(min(a, b), c) + 1;
  if constexpr (OP == OP_DP4A)         asm volatile("dp4a.u32.u32 %0, %1, %2, %0;" : "+r"(a) : "r"(b), "r"(c));
  if constexpr (OP == OP_SHF)          asm volatile("shf.r.wrap.b32 %0, %0, %1, 8;" : "+r"(a) : "r"(b));
  if constexpr (OP == OP_MIX_IADD3_IMAD) {
    asm volatile("add.u32 %0, %0, %1;\n\tadd.u32 %0, %0, %2;" : "+r"(a) : "r"(b), "r"(c));
    asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(a) : "r"(b), "r"(c));
  }
  if constexpr (OP == OP_MIX_VABS_IMAD) {
    asm volatile("vabsdiff4.u32.u32.u32 %0, %0, %1, %2;" : "+r"(a) : "r"(b), "r"(c));
    asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(a) : "r"(b), "r"(c));
  }
  if constexpr (OP == OP_MIX_PRMT_IMAD) {
    asm volatile("prmt.b32 %0, %0, %1, 0x5140;" : "+r"(a) : "r"(b));
    asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(a) : "r"(b), "r"(c));
  }
  if constexpr (OP == OP_MIX_VIMNMX_IMAD) {
    asm volatile("min.u32 %0, %0, %1;" : "+r"(a) : "r"(b));
    asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(a) : "r"(b), "r"(c));
  }
  if constexpr (OP == OP_LDS128) {
    uint4 v; asm volatile("ld.volatile.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr) : "memory");
    a ^= v.x;  // one LOP3 per load rides along; counted as 2 instructions
  }
  if constexpr (OP == OP_MIX_LDS128_IADD3) {
    uint4 v; asm volatile("ld.volatile.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr) : "memory");
    asm volatile("add.u32 %0, %0, %1;\n\tadd.u32 %0, %0, %2;" : "+r"(a) : "r"(v.x), "r"(v.y));
    asm volatile("add.u32 %0, %0, %1;\n\tadd.u32 %0, %0, %2;" : "+r"(a) : "r"(v.z), "r"(v.w));
    asm volatile("add.u32 %0, %0, %1;\n\tadd.u32 %0, %0, %2;" : "+r"(a) : "r"(b), "r"(c));
    asm volatile("add.u32 %0, %0, %1;\n\tadd.u32 %0, %0, %2;" : "+r"(a) : "r"(c), "r"(b));
  }
  if constexpr (OP == OP_REDUX_MIN)    asm volatile("redux.sync.min.u32 %0, %0, 0xffffffff;" : "+r"(a));
  if constexpr (OP == OP_SHFL_BFLY)    asm volatile("shfl.sync.bfly.b32 %0, %0, 1, 0x1f, 0xffffffff;" : "+r"(a));
  if constexpr (OP == OP_MIX_REDUX_8ALU) {
    uint32_t r; asm volatile("redux.sync.min.u32 %0, %1, 0xffffffff;" : "=r"(r) : "r"(a));
    asm volatile("add.u32 %0, %0, %1;\n\tadd.u32 %0, %0, %2;" : "+r"(a) : "r"(r), "r"(c));
    asm volatile("add.u32 %0, %0, %1;\n\tadd.u32 %0, %0, %2;" : "+r"(a) : "r"(b), "r"(c));
    asm volatile("add.u32 %0, %0, %1;\n\tadd.u32 %0, %0, %2;" : "+r"(a) : "r"(c), "r"(b));
    asm volatile("add.u32 %0, %0, %1;\n\tadd.u32 %0, %0, %2;" : "+r"(a) : "r"(b), "r"(c));
    asm volatile("add.u32 %0, %0, %1;\n\tadd.u32 %0, %0, %2;" : "+r"(a) : "r"(c), "r"(b));
    asm volatile("add.u32 %0, %0, %1;\n\tadd.u32 %0, %0, %2;" : "+r"(a) : "r"(b), "r"(c));
    asm volatile("add.u32 %0, %0, %1;\n\tadd.u32 %0, %0, %2;" : "+r"(a) : "r"(c), "r"(b));
    asm volatile("add.u32 %0, %0, %1;\n\tadd.u32 %0, %0, %2;" : "+r"(a) : "r"(b), "r"(c));
  }
  if constexpr (OP == OP_MIX_SHFL_8ALU) {
    uint32_t r; asm volatile("shfl.sync.bfly.b32 %0, %1, 1, 0x1f, 0xffffffff;" : "=r"(r) : "r"(a));
    asm volatile("add.u32 %0, %0, %1;\n\tadd.u32 %0, %0, %2;" : "+r"(a) : "r"(r), "r"(c));
    asm volatile("add.u32 %0, %0, %1;\n\tadd.u32 %0, %0, %2;" : "+r"(a) : "r"(b), "r"(c));
    asm volatile("add.u32 %0, %0, %1;\n\tadd.u32 %0, %0, %2;" : "+r"(a) : "r"(c), "r"(b));
    asm volatile("add.u32 %0, %0, %1;\n\tadd.u32 %0, %0, %2;" : "+r"(a) : "r"(b), "r"(c));
    asm volatile("add.u32 %0, %0, %1;\n\tadd.u32 %0, %0, %2;" : "+r"(a) : "r"(c), "r"(b));
    asm volatile("add.u32 %0, %0, %1;\n\tadd.u32 %0, %0, %2;" : "+r"(a) : "r"(b), "r"(c));
    asm volatile("add.u32 %0, %0, %1;\n\tadd.u32 %0, %0, %2;" : "+r"(a) : "r"(c), "r"(b));
    asm volatile("add.u32 %0, %0, %1;\n\tadd.u32 %0, %0, %2;" : "+r"(a) : "r"(b), "r"(c));
  }
  if constexpr (OP == OP_MIX_ALU3_IMAD1) {
    asm volatile("vabsdiff4.u32.u32.u32 %0, %0, %1, %2;" : "+r"(a) : "r"(b), "r"(c));
    asm volatile("prmt.b32 %0, %0, %1, 0x5140;" : "+r"(a) : "r"(b));
    asm volatile("add.u32 %0, %0, %1;\n\tadd.u32 %0, %0, %2;" : "+r"(a) : "r"(b), "r"(c));
    asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(a) : "r"(b), "r"(c));
  }
}

template <int OP>
__global__ void __launch_bounds__(1024, 1) bench_kernel(uint32_t* out, long long* cycles, uint32_t seed) {
  __shared__ uint4 sm[1024];
  sm[threadIdx.x] = make_uint4(threadIdx.x, seed, threadIdx.x ^ seed, 1);
  __syncthreads();
  uint32_t addr = (uint32_t)__cvta_generic_to_shared(&sm[threadIdx.x]);
  uint32_t v[CHAINS];
#pragma unroll
  for (int i = 0; i < CHAINS; ++i) v[i] = threadIdx.x * 2654435761u + i * seed;
  uint32_t b = seed | 1u, c = seed * 3u + threadIdx.x;
  long long t0 = clock64();
#pragma unroll 1
  for (int it = 0; it < ITERS; ++it) {
#pragma unroll
    for (int u = 0; u < 4; ++u) {
#pragma unroll
      for (int i = 0; i < CHAINS; ++i) step<OP>(v[i], b, c, sm, addr);
    }
  }
  long long t1 = clock64();
  uint32_t acc = 0;
#pragma unroll
  for (int i = 0; i < CHAINS; ++i) acc ^= v[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
  if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

template <int OP>
static void run(int sms, uint32_t* d_out, long long* d_cyc, bool first) {
  const int grid = sms, block = 1024;
  bench_kernel<OP><<<grid, block>>>(d_out, d_cyc, 12345u);   // warm-up
  CK(cudaDeviceSynchronize());
  cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  CK(cudaEventRecord(e0));
  const int reps = 5;
  for (int r = 0; r < reps; ++r) bench_kernel<OP><<<grid, block>>>(d_out, d_cyc, 777u + r);
  CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
  float ms = 0; CK(cudaEventElapsedTime(&ms, e0, e1));
  long long* h = (long long*)malloc(sizeof(long long) * grid);
  CK(cudaMemcpy(h, d_cyc, sizeof(long long) * grid, cudaMemcpyDeviceToHost));
  double cyc = 0; for (int i = 0; i < grid; ++i) cyc += (double)h[i]; cyc /= grid;
  free(h);
  const double instr_per_thread = (double)ITERS * 4 * CHAINS * op_instr[OP];
  const double lane_ops_per_sm = instr_per_thread * block;           // one CTA per SM
  const double per_clk_sm = lane_ops_per_sm / cyc;
  const double tops = lane_ops_per_sm * grid * reps / (ms * 1e-3) / 1e12;
  printf("%s  \"%s\": {\"lane_ops_per_clk_per_sm\": %.2f, \"tera_lane_ops_per_s\": %.3f, \"cycles\": %.0f, \"ms\": %.4f}",
         first ? "" : ",\n", op_name[OP], per_clk_sm, tops, cyc, ms / reps);
}

template <int OP> struct RunAll { static void go(int sms, uint32_t* o, long long* c) { run<OP>(sms, o, c, OP == 0); RunAll<OP + 1>::go(sms, o, c); } };
template <> struct RunAll<OP_COUNT> { static void go(int, uint32_t*, long long*) {} };

int main() {
  cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, 0));
  int clk_khz = 0; CK(cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, 0));
  uint32_t* d_out; long long* d_cyc;
  CK(cudaMalloc(&d_out, sizeof(uint32_t) * p.multiProcessorCount * 1024));
  CK(cudaMalloc(&d_cyc, sizeof(long long) * p.multiProcessorCount));
  printf("{\n  \"device\": \"%s\", \"sms\": %d, \"cc\": \"%d.%d\", \"clock_rate_khz\": %d, \"smem_per_sm\": %zu, \"smem_optin\": %zu, \"regs_per_sm\": %d, \"l2_bytes\": %d,\n  \"note\": \"lane-ops = 32 x warp instructions; mixes count every instruction; 1 CTA x 1024 threads per SM\",\n",
         p.name, p.multiProcessorCount, p.major, p.minor, clk_khz, p.sharedMemPerMultiprocessor, p.sharedMemPerBlockOptin, p.regsPerMultiprocessor, p.l2CacheSize);
  RunAll<0>::go(p.multiProcessorCount, d_out, d_cyc);
  printf("\n}\n");
  return 0;
}
