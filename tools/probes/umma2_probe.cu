// Micro-benchmark of tcgen05.mma.cta_group::2 (CTA pair, M = 256) issue/execution rate on sm_100a, shared-memory operands
// (not product code; evidence for DESIGN.md). One cluster of two CTAs per SM pair; the leader CTA issues; operands are
// zero-filled SW128 K-major tiles (A: 128 rows per CTA, B: N/2 rows per CTA), no loads.
// Compare with umma_probe.cu (cta_group::1): cycles per MMA of the same per-CTA work (128 x N x 16).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o umma2_probe umma2_probe.cu ; run: ./umma2_probe
#include <cstdio>
#include <cstdlib>
#include "../../can-image-style-transfer-save-automotive-radar_b200/csrc/ptx.cuh"
using namespace ist;

template <uint32_t kCols>
__device__ __forceinline__ void tmem_alloc2(uint32_t smem_dst) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_dst), "n"(kCols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
template <uint32_t kCols>
__device__ __forceinline__ void tmem_dealloc2(uint32_t taddr) {
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "n"(kCols) : "memory");
}
__device__ __forceinline__ void umma2_f16_lh(uint32_t d_tmem, uint32_t a_lo, uint32_t a_hi, uint32_t b_lo, uint32_t b_hi, uint32_t idesc, uint32_t acc) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t.reg .b64 da, db;\n\tmov.b64 da, {%1, %2};\n\tmov.b64 db, {%3, %4};\n\tsetp.ne.b32 p, %6, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], da, db, %5, p;\n\t}\n" ::"r"(d_tmem), "r"(a_lo), "r"(a_hi), "r"(b_lo), "r"(b_hi), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void umma2_commit(uint32_t bar, uint16_t mask) {
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(bar), "h"(mask) : "memory");
}

__device__ __forceinline__ void umma2_commit_local(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
struct Args { int n, reps, cper, issuers, mc; long long* out; };

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(128, 1) probe2(Args a) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const uint32_t sA = base, sB = base + 32768, bar = base + 32768 + 65536, slot = bar + 128;
    volatile uint32_t* slot_g = reinterpret_cast<volatile uint32_t*>(smem_raw + (slot - smem_u32(smem_raw)));
    for (uint32_t i = threadIdx.x; i < (32768 + 65536) / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem_raw + (base - smem_u32(smem_raw)))[i] = 0;
    const int warp = threadIdx.x >> 5;
    const uint32_t rank = cluster_ctarank();
    if (threadIdx.x == 0) { for (int i = 0; i < 16; ++i) mbar_init(bar + 8 * i, 1); fence_barrier_init(); }
    if (warp == 0) tmem_alloc2<512>(slot);
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    tc_fence_before(); cluster_sync_all(); tc_fence_after();
    const uint32_t tmem = *slot_g;
    const uint32_t idesc = umma_idesc_f16(UMMA_FMT_F16, 256, a.n, 0, 0);
    const uint32_t hi_w = (1024u >> 4) | (1u << 14) | (2u << 29);
    if (rank == 0 && warp < a.issuers) {
        long long t0 = 0;
        if (elect_one()) {
            t0 = clock64();
            const uint32_t d = tmem + (uint32_t)warp * (a.n > 128 ? 256u : 128u);
            for (int r = 0; r < a.reps; ++r) {
                const uint32_t k4 = r & 3;
                umma2_f16_lh(d, (sA >> 4) + 2 * k4 + (uint32_t)((r >> 2) & 7) * 8, hi_w, (sB >> 4) + 2 * k4, hi_w, idesc, 1u);
                // cper is a power of two (mask test: no integer division in the issuing thread); mc = 1: multicast to both CTAs,
                // mc = 0: plain cta_group::2 commit to the leader's barrier
                if (a.cper > 0 && (r & (a.cper - 1)) == a.cper - 1 && r != a.reps - 1) {
                    if (a.mc) umma2_commit(bar + 64 + 8 * warp, 3); else umma2_commit_local(bar + 64 + 8 * warp);
                }
            }
            umma2_commit(bar + 8 * warp, 1);
            const long long t1 = clock64();
            mbar_wait(bar + 8 * warp, 0);
            const long long t2 = clock64();
            a.out[((blockIdx.x >> 1) * 4 + warp) * 2 + 0] = t1 - t0;
            a.out[((blockIdx.x >> 1) * 4 + warp) * 2 + 1] = t2 - t0;
        }
        __syncwarp();
    }
    tc_fence_before(); cluster_sync_all();
    if (warp == 0) { tc_fence_after(); tmem_dealloc2<512>(tmem); }
}

int main() {
    long long* out; cudaMalloc(&out, sizeof(long long) * 74 * 8);
    const int smem = 32768 + 65536 + 1024 + 512;
    cudaFuncSetAttribute(probe2, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    printf("cta_group::2, M = 256 (128 rows per CTA); cycles per MMA (per-CTA work 128 x N x 16; floor N/2)\n");
    printf("%4s %5s %2s %8s | %13s %13s\n", "N", "cper", "mc", "issuers", "issue clk/MMA", "total clk/MMA");
    const int reps = 4096;
    for (int n : {64, 128, 256})
        for (int cper : {0, 1, 4, 8})
          for (int mc = 0; mc < 2; ++mc)
            for (int issuers = 1; issuers <= 2; ++issuers) {
                if (cper == 0 && mc == 1) continue;
                Args a{n, reps, cper, issuers, mc, out};
                cudaMemset(out, 0, sizeof(long long) * 74 * 8);
                probe2<<<148, 128, smem>>>(a);
                cudaError_t e = cudaDeviceSynchronize();
                if (e != cudaSuccess) { printf("N %d: %s\n", n, cudaGetErrorString(e)); return 1; }
                long long h[74 * 8]; cudaMemcpy(h, out, sizeof(h), cudaMemcpyDeviceToHost);
                double is = 0, tot = 0;
                for (int c = 0; c < 74; ++c) {
                    long long mi = 0, mt = 0;
                    for (int w = 0; w < issuers; ++w) { if (h[c * 8 + 2 * w] > mi) mi = h[c * 8 + 2 * w]; if (h[c * 8 + 2 * w + 1] > mt) mt = h[c * 8 + 2 * w + 1]; }
                    is += mi; tot += mt;
                }
                printf("%4d %5d %2d %8d | %13.1f %13.1f\n", n, cper, mc, issuers, is / 74 / reps / issuers, tot / 74 / reps / issuers);
            }
    return 0;
}
