// Micro-benchmark of tcgen05.mma issue/execution rate on sm_100a (not product code; evidence for DESIGN.md).
// One CTA per SM, operands are zero-filled shared memory (SW128 K-major tiles), no loads: measures cycles per MMA for
//   N in {64,128,256}, A from shared memory (SS) or from tensor memory (TS), commit every `cper` MMAs, 1 or 2 issuing warps.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o umma_probe umma_probe.cu ; run: ./umma_probe
#include <cstdio>
#include <cstdlib>
#include "../../can-image-style-transfer-save-automotive-radar_b200/csrc/ptx.cuh"
using namespace ist;

__device__ __forceinline__ void umma_f16_ts(uint32_t d_tmem, uint32_t a_tmem, uint32_t b_lo, uint32_t b_hi, uint32_t idesc, uint32_t acc) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t.reg .b64 db;\n\tmov.b64 db, {%2, %3};\n\tsetp.ne.b32 p, %5, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], db, %4, p;\n\t}\n" ::"r"(d_tmem), "r"(a_tmem), "r"(b_lo), "r"(b_hi), "r"(idesc), "r"(acc) : "memory");
}

__device__ __forceinline__ void utccp_128x256b(uint32_t taddr, uint32_t lo, uint32_t hi) {
    asm volatile("{\n\t.reg .b64 d;\n\tmov.b64 d, {%1, %2};\n\ttcgen05.cp.cta_group::1.128x256b [%0], d;\n\t}\n" ::"r"(taddr), "r"(lo), "r"(hi) : "memory");
}
struct Args { int n, reps, cper, issuers, ts, nacc; long long* out; };

__global__ void __launch_bounds__(128, 1) probe(Args a) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const uint32_t sA = base, sB = base + 32768, bar = base + 32768 + 65536, slot = bar + 64;
    volatile uint32_t* slot_g = reinterpret_cast<volatile uint32_t*>(smem_raw + (slot - smem_u32(smem_raw)));
    for (uint32_t i = threadIdx.x; i < (32768 + 65536) / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem_raw + (base - smem_u32(smem_raw)))[i] = 0;
    const int warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) { for (int i = 0; i < 8; ++i) mbar_init(bar + 8 * i, 1); fence_barrier_init(); }
    if (warp == 0) tmem_alloc<512>(slot);
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    tc_fence_before(); __syncthreads(); tc_fence_after();
    const uint32_t tmem = *slot_g;
    const uint32_t idesc = umma_idesc_f16(UMMA_FMT_F16, 128, a.n, 0, 0);
    const uint32_t hi_w = (1024u >> 4) | (1u << 14) | (2u << 29);
    if (warp < a.issuers) {
        long long t0 = 0;
        if (elect_one()) {
            t0 = clock64();
            const uint32_t dbase = tmem + (uint32_t)warp * 128u;
            for (int r = 0; r < a.reps; ++r) {
                const uint32_t d = (a.n > 128) ? (tmem + (uint32_t)warp * 256u) : (a.ts == 2 ? tmem + (uint32_t)warp * 128u + (uint32_t)(((r % 3) != 0) ? 64u * (a.n <= 64) : 0u) : dbase);
                const uint32_t k4 = r & 3;
                if (a.ts == 2) {
                    // conv pattern: every 3 MMAs (hi*hi, hi*lo, lo*hi of one k16 slice) are preceded by 2 tcgen05.cp (A_hi, A_lo slices);
                    // staging ring of 4 slice pairs per issuer in TMEM columns [384 + 64*warp, +64)
                    const uint32_t g = (uint32_t)(r / 3), stg = tmem + 384 + 64 * warp + (g & 3) * 16;
                    if (r % 3 == 0) {
                        utccp_128x256b(stg, (sA >> 4) + 2 * (g & 3) + (uint32_t)((g >> 2) & 7) * 8, hi_w);
                        utccp_128x256b(stg + 8, (sA >> 4) + 1024 + 2 * (g & 3), hi_w);
                    }
                    umma_f16_ts(d, stg + ((r % 3) == 2 ? 8 : 0), (sB >> 4) + 2 * (g & 3) + ((r % 3) == 1 ? 1024 : 0), hi_w, idesc, 1u);
                } else if (a.ts) umma_f16_ts(d, tmem + 480 + 0, (sB >> 4) + 2 * k4, hi_w, idesc, 1u);
                else umma_f16_lh(d, (sA >> 4) + 2 * k4 + (uint32_t)((r >> 2) & 7) * 8, hi_w, (sB >> 4) + 2 * k4, hi_w, idesc, 1u);
                if (a.cper > 0 && (r % a.cper) == a.cper - 1 && r != a.reps - 1) umma_commit(bar + 32 + 8 * warp);   // dummy barrier, never waited
            }
            umma_commit(bar + 8 * warp);
            const long long t1 = clock64();
            mbar_wait(bar + 8 * warp, 0);
            const long long t2 = clock64();
            a.out[(blockIdx.x * 4 + warp) * 2 + 0] = t1 - t0;
            a.out[(blockIdx.x * 4 + warp) * 2 + 1] = t2 - t0;
        }
        __syncwarp();
    }
    tc_fence_before(); __syncthreads();
    if (warp == 0) { tc_fence_after(); tmem_dealloc<512>(tmem); }
}

int main() {
    long long* out; cudaMalloc(&out, sizeof(long long) * 148 * 8);
    const int smem = 32768 + 65536 + 1024 + 256;
    cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    printf("%4s %3s %5s %8s %5s | %10s %10s\n", "N", "TS", "cper", "issuers", "nacc", "issue clk/MMA", "total clk/MMA");
    const int reps = 4096;
    for (int ts = 0; ts < 3; ++ts)
        for (int n : {64, 128, 256})
            for (int cper : {0, 4, 12})
                for (int issuers = 1; issuers <= 4; ++issuers)
                    for (int nacc : {1}) {
                        if (n == 256 && issuers > 2) continue;
                        if (ts == 2 && (n == 256 || issuers > 2 || cper == 4)) continue;
                        if (ts && cper == 12) continue;
                        Args a{n, reps, cper, issuers, ts, nacc, out};
                        cudaMemset(out, 0, sizeof(long long) * 148 * 8);
                        probe<<<148, 128, smem>>>(a);
                        cudaError_t e = cudaDeviceSynchronize();
                        if (e != cudaSuccess) { printf("N %d ts %d: %s\n", n, ts, cudaGetErrorString(e)); return 1; }
                        long long h[148 * 8]; cudaMemcpy(h, out, sizeof(h), cudaMemcpyDeviceToHost);
                        double is = 0, tot = 0;
                        for (int c = 0; c < 148; ++c) {
                            long long mi = 0, mt = 0;
                            for (int w = 0; w < issuers; ++w) { if (h[c * 8 + 2 * w] > mi) mi = h[c * 8 + 2 * w]; if (h[c * 8 + 2 * w + 1] > mt) mt = h[c * 8 + 2 * w + 1]; }
                            is += mi; tot += mt;
                        }
                        // cycles per MMA over ALL issuers (issuers * reps MMAs in `tot` cycles)
                        printf("%4d %3d %5d %8d %5d | %10.1f %10.1f\n", n, ts, cper, issuers, nacc, is / 148 / reps / issuers, tot / 148 / reps / issuers);
                    }
    return 0;
}
