// Hardware probe (bring-up tool, not product code): which shared-memory operand descriptors does tcgen05.mma accept?
//  A) 128B-swizzled K-major B operand whose start address is shifted by s pixel rows (s * 128 B) inside a larger
//     swizzled buffer -- with the descriptor's base_offset field = 0 or = (addr >> 7) & 7.  Needed for 3x3 convolutions
//     that address all nine taps inside ONE halo tile kept in shared memory.
//  B) no-swizzle K-major B operand with LBO = 16 B, SBO = 128 B, i.e. overlapping 8 x 16 B core matrices: row n of
//     K chunk j sits at base + 16 * (n + j).  Needed for the stride-2 7x7 stem's implicit im2col from a parity-split row.
// One CTA, one thread issues 4 (A) / 1 (B) MMAs of M = 128, N = 128; results are compared on the host.
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(1); } } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }

__device__ __forceinline__ uint64_t make_desc(uint32_t addr, uint32_t lbo, uint32_t sbo, uint32_t layout, uint32_t base_off) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((addr & 0x3FFFF) >> 4);
  d |= static_cast<uint64_t>((lbo >> 4) & 0x3FFF) << 16;
  d |= static_cast<uint64_t>((sbo >> 4) & 0x3FFF) << 32;
  d |= static_cast<uint64_t>(1) << 46;
  d |= static_cast<uint64_t>(base_off & 7) << 49;
  d |= static_cast<uint64_t>(layout) << 61;
  return d;
}

struct Args {
  int mode;       // 0 = swizzled + row shift, 1 = no-swizzle overlapping
  int shift;      // rows (mode 0) / 16-byte pixels (mode 1)
  int base_mode;  // mode 0: 0 -> base_offset 0, 1 -> (addr >> 7) & 7
  int n;          // MMA N
};

__global__ void __launch_bounds__(128, 1) probe(const __nv_bfloat16* a, const __nv_bfloat16* bbuf, int brows, float* out, Args g) {
  extern __shared__ uint8_t raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sA = smem;                    // 128 x 64 bf16 swizzled: 16 KB
  uint8_t* sB = smem + 16384;            // brows x 128 B
  __shared__ uint64_t bar;
  __shared__ uint32_t tslot;
  const int tid = threadIdx.x, warp = tid >> 5;
  for (int i = tid; i < 128 * 64; i += 128) {
    const int m = i / 64, k = i % 64;
    *reinterpret_cast<__nv_bfloat16*>(sA + m * 128 + (((k >> 3) ^ (m & 7)) << 4) + (k & 7) * 2) = a[i];
  }
  if (g.mode == 0) {
    for (int i = tid; i < brows * 64; i += 128) {
      const int r = i / 64, c = i % 64;
      *reinterpret_cast<__nv_bfloat16*>(sB + r * 128 + (((c >> 3) ^ (r & 7)) << 4) + (c & 7) * 2) = bbuf[i];
    }
  } else {
    for (int i = tid; i < brows * 64; i += 128) reinterpret_cast<__nv_bfloat16*>(sB)[i] = bbuf[i];   // linear
  }
  if (tid == 0) {
    uint32_t ba = smem_u32(&bar);
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(ba));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tslot)), "r"(256u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = tslot;
  if (tid == 0) {
    const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | (static_cast<uint32_t>(g.n >> 3) << 17) | (static_cast<uint32_t>(128 >> 4) << 24);
    const uint32_t aaddr = smem_u32(sA);
    const int ksteps = g.mode == 0 ? 4 : 1;
    for (int k = 0; k < ksteps; ++k) {
      const uint64_t da = make_desc(aaddr + k * 32, 16, 1024, 2, 0);
      uint64_t db;
      if (g.mode == 0) {
        const uint32_t baddr = smem_u32(sB) + g.shift * 128 + k * 32;
        db = make_desc(baddr, 16, 1024, 2, g.base_mode ? ((baddr >> 7) & 7) : 0);
      } else {
        const uint32_t baddr = smem_u32(sB) + g.shift * 16;
        db = make_desc(baddr, 16, 128, 0, 0);
      }
      const uint32_t acc = k > 0;
      asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem), "l"(da), "l"(db), "r"(idesc), "r"(acc) : "memory");
    }
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
  }
  // wait
  {
    uint32_t done = 0, spins = 0;
    while (!done) {
      asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.b32 %0, 1, 0, p;\n\t}" : "=r"(done) : "r"(smem_u32(&bar)), "r"(0u) : "memory");
      if (!done && ++spins > (1u << 24)) { if (tid == 0) printf("timeout\n"); break; }
    }
  }
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  for (int c0 = 0; c0 < g.n; c0 += 32) {
    uint32_t r[32];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(tmem + ((static_cast<uint32_t>(warp * 32)) << 16) + c0) : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    for (int j = 0; j < 32; ++j) out[tid * 256 + c0 + j] = __uint_as_float(r[j]);
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(256u) : "memory");
}

// mode 2: issue-rate / operand-fetch cost of back-to-back MMAs from shared memory: cycles per tcgen05.mma for (M, N)
__global__ void __launch_bounds__(128, 1) rate(int m, int n, int iters, int bshift_rows, long long* cycles) {
  extern __shared__ uint8_t raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(raw) + 1023) & ~uintptr_t(1023));
  __shared__ uint64_t bar;
  __shared__ uint32_t tslot;
  const int tid = threadIdx.x, warp = tid >> 5;
  for (int i = tid; i < (16384 + 65536) / 4; i += 128) reinterpret_cast<uint32_t*>(smem)[i] = 0;
  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tslot)), "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = tslot;
  if (tid == 0) {
    const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | (static_cast<uint32_t>(n >> 3) << 17) | (static_cast<uint32_t>(m >> 4) << 24);
    const uint32_t aaddr = smem_u32(smem), baddr = smem_u32(smem + 16384);
    const long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
      const int k = i & 3;
      const uint64_t da = make_desc(aaddr + k * 32, 16, 1024, 2, 0);
      const uint64_t db = make_desc(baddr + ((i >> 2) % 9) * bshift_rows * 128 + k * 32, 16, 1024, 2, 0);
      asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem + (i & 1) * 256), "l"(da), "l"(db), "r"(idesc), "r"(1u) : "memory");
    }
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
    uint32_t done = 0;
    while (!done) asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.b32 %0, 1, 0, p;\n\t}" : "=r"(done) : "r"(smem_u32(&bar)), "r"(0u) : "memory");
    *cycles = clock64() - t0;
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u) : "memory");
}

int main(int argc, char** argv) {
  if (argc > 1) {   // timing mode
    long long* dc; CK(cudaMalloc(&dc, 8));
    const int smem = 16384 + 65536 + 2048 + 32768;
    CK(cudaFuncSetAttribute(rate, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    const int mn[][2] = {{128, 256}, {128, 128}, {128, 64}, {128, 32}, {64, 256}, {64, 128}, {64, 64}, {128, 192}, {128, 96}, {128, 16}};
    for (auto& p : mn)
      for (int sh : {0, 1}) {
        long long hc = 0;
        rate<<<1, 128, smem>>>(p[0], p[1], 64, sh, dc); CK(cudaDeviceSynchronize());   // warm-up
        rate<<<1, 128, smem>>>(p[0], p[1], 4096, sh, dc); CK(cudaDeviceSynchronize());
        CK(cudaMemcpy(&hc, dc, 8, cudaMemcpyDeviceToHost));
        printf("M %3d N %3d b-shift %d rows: %.1f cycles per MMA (K = 16) -> %.0f %% of the 8192 FLOP/cycle/SM peak\n", p[0], p[1], sh,
               hc / 4096.0, 100.0 * (2.0 * p[0] * p[1] * 16) / (hc / 4096.0) / 8192.0);
      }
    return 0;
  }
  const int brows = 512;
  std::vector<__nv_bfloat16> ha(128 * 64), hb(brows * 64);
  std::vector<float> fa(128 * 64), fb(brows * 64);
  srand(1);
  for (int i = 0; i < 128 * 64; ++i) { fa[i] = static_cast<float>((rand() % 7) - 3); ha[i] = __float2bfloat16(fa[i]); }
  for (int i = 0; i < brows * 64; ++i) { fb[i] = static_cast<float>((rand() % 9) - 4); hb[i] = __float2bfloat16(fb[i]); }
  __nv_bfloat16 *da, *db; float* dout;
  CK(cudaMalloc(&da, ha.size() * 2)); CK(cudaMalloc(&db, hb.size() * 2)); CK(cudaMalloc(&dout, 128 * 256 * 4));
  CK(cudaMemcpy(da, ha.data(), ha.size() * 2, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(db, hb.data(), hb.size() * 2, cudaMemcpyHostToDevice));
  const int smem = 16384 + brows * 128 + 2048;
  CK(cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  std::vector<float> ho(128 * 256);
  auto run = [&](Args g) {
    CK(cudaMemset(dout, 0, 128 * 256 * 4));
    probe<<<1, 128, smem>>>(da, db, brows, dout, g);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("mode %d shift %d base %d n %d: LAUNCH ERROR %s\n", g.mode, g.shift, g.base_mode, g.n, cudaGetErrorString(e)); exit(2); }
    CK(cudaMemcpy(ho.data(), dout, ho.size() * 4, cudaMemcpyDeviceToHost));
    int bad = 0; double maxerr = 0;
    for (int m = 0; m < 128; ++m)
      for (int n = 0; n < g.n; ++n) {
        double ref = 0;
        if (g.mode == 0) { for (int k = 0; k < 64; ++k) ref += fa[m * 64 + k] * fb[(g.shift + n) * 64 + k]; }
        else { for (int k = 0; k < 16; ++k) ref += fa[m * 64 + k] * fb[(n + g.shift + k / 8) * 8 + k % 8]; }
        const double err = fabs(ref - ho[m * 256 + n]);
        if (err > 1e-3) ++bad;
        if (err > maxerr) maxerr = err;
      }
    printf("mode %d shift %3d base_mode %d N %3d : %s (bad %d, max err %.1f)\n", g.mode, g.shift, g.base_mode, g.n, bad ? "FAIL" : "PASS", bad, maxerr);
  };
  for (int n : {128, 256})
    for (int s : {0, 8, 1, 3, 7, 9, 66, 67, 133})
      for (int bm : {0, 1}) run(Args{0, s, bm, n});
  for (int n : {128, 256})
    for (int s : {0, 1, 5}) run(Args{1, s, 0, n});
  return 0;
}
