// Probe: does tcgen05.cp.128x256b with a SW128 K-major descriptor (arbitrary 128-byte-row start, SBO = 10 rows: the halo
// addressing of conv_halo.cuh) deliver the A operand of a TS-mode tcgen05.mma?  (not product code)
//  1. fill a halo-like smem buffer (180 rows x 64 fp16, TMA SWIZZLE_128B layout) with known values
//  2. tcgen05.cp 128x256b -> TMEM columns, read back with tcgen05.ld, compare with the expected rows / channels
//  3. D_ts = A(TMEM) x B(smem) vs D_ss = A(smem) x B(smem), compare
#include <cstdio>
#include <cuda_fp16.h>
#include "../../can-image-style-transfer-save-automotive-radar_b200/csrc/ptx.cuh"
using namespace ist;

__device__ __forceinline__ void utccp_128x256b(uint32_t taddr, uint32_t lo, uint32_t hi) {
    asm volatile("{\n\t.reg .b64 d;\n\tmov.b64 d, {%1, %2};\n\ttcgen05.cp.cta_group::1.128x256b [%0], d;\n\t}\n" ::"r"(taddr), "r"(lo), "r"(hi) : "memory");
}
__device__ __forceinline__ void umma_ts(uint32_t d_tmem, uint32_t a_tmem, uint32_t b_lo, uint32_t b_hi, uint32_t idesc, uint32_t acc) {
    asm volatile("{\n\t.reg .pred p;\n\t.reg .b64 db;\n\tmov.b64 db, {%2, %3};\n\tsetp.ne.b32 p, %5, 0;\n\t"
                 "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], db, %4, p;\n\t}\n" ::"r"(d_tmem), "r"(a_tmem), "r"(b_lo), "r"(b_hi), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void tmem_ld_32x32_x8(uint32_t taddr, uint32_t (&r)[8]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]) : "r"(taddr) : "memory");
}

__global__ void __launch_bounds__(128, 1) probe(int rowoff, int k4, int* out_mismatch, float* out_d /*[2][128][64]*/, uint32_t* out_raw) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    uint8_t* gbase = smem_raw + (base - smem_u32(smem_raw));
    const uint32_t sA = base, sB = base + 24576, bar = base + 24576 + 8192, slot = bar + 64;
    volatile uint32_t* slot_g = reinterpret_cast<volatile uint32_t*>(gbase + 24576 + 8192 + 64);
    // A: 180 rows x 64 ch, value(r, c) = (r * 7 + c) % 512 - 100   (exact in fp16)
    for (int i = threadIdx.x; i < 180 * 64; i += blockDim.x) {
        const int r = i / 64, c = i % 64;
        const int off = r * 128 + (((c >> 3) ^ (r & 7)) << 4) + (c & 7) * 2;
        *reinterpret_cast<__half*>(gbase + off) = __float2half((float)((r * 7 + c) % 512 - 100));
    }
    // B: 64 rows (n) x 64 ch K-major SW128, value(n, c) = ((n * 3 + c * 5) % 17) - 8
    for (int i = threadIdx.x; i < 64 * 64; i += blockDim.x) {
        const int n = i / 64, c = i % 64;
        const int off = n * 128 + (((c >> 3) ^ (n & 7)) << 4) + (c & 7) * 2;
        *reinterpret_cast<__half*>(gbase + 24576 + off) = __float2half((float)((n * 3 + c * 5) % 17 - 8));
    }
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) { mbar_init(bar, 1); mbar_init(bar + 8, 1); fence_barrier_init(); }
    if (warp == 0) tmem_alloc<256>(slot);
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    tc_fence_before(); __syncthreads(); tc_fence_after();
    const uint32_t tmem = *slot_g;
    const uint32_t a_hi_w = ((1280u >> 4) & 0x3FFFu) | (1u << 14) | (2u << 29);
    const uint32_t b_hi_w = (1024u >> 4) | (1u << 14) | (2u << 29);
    const uint32_t a_lo = ((sA + rowoff * 128) >> 4) + 2 * k4, b_lo = (sB >> 4) + 2 * k4;
    const uint32_t idesc = umma_idesc_f16(UMMA_FMT_F16, 128, 64, 0, 0);
    if (warp == 0) {
        if (elect_one()) {
            utccp_128x256b(tmem + 192, a_lo, a_hi_w);                    // A slice -> TMEM columns 192..199
            umma_ts(tmem + 0, tmem + 192, b_lo, b_hi_w, idesc, 0u);      // D_ts  -> columns 0..63
            umma_f16_lh(tmem + 64, a_lo, a_hi_w, b_lo, b_hi_w, idesc, 0u);   // D_ss -> columns 64..127
            umma_commit(bar);
        }
        __syncwarp();
    }
    mbar_wait(bar, 0);
    tc_fence_after();
    const int m = warp * 32 + lane;
    const uint32_t lane_base = tmem + ((uint32_t)(warp * 32) << 16);
    uint32_t r8[8];
    tmem_ld_32x32_x8(lane_base + 192, r8);
    tmem_ld_wait();
    int bad = 0;
    const int hr = rowoff + (m / 8) * 10 + (m % 8);
    for (int j = 0; j < 8; ++j) {
        out_raw[m * 8 + j] = r8[j];
        for (int e = 0; e < 2; ++e) {
            const int c = k4 * 16 + 2 * j + e;
            const float want = (float)((hr * 7 + c) % 512 - 100);
            const float got = __half2float(__ushort_as_half((unsigned short)((r8[j] >> (16 * e)) & 0xFFFF)));
            if (want != got) ++bad;
        }
    }
    atomicAdd(out_mismatch, bad);
    for (int half = 0; half < 2; ++half)
        for (int c0 = 0; c0 < 64; c0 += 32) {
            uint32_t r[32];
            tmem_ld_32x32(lane_base + half * 64 + c0, r);
            tmem_ld_wait();
            for (int j = 0; j < 32; ++j) out_d[(half * 128 + m) * 64 + c0 + j] = __uint_as_float(r[j]);
        }
    tc_fence_before(); __syncthreads();
    if (warp == 0) { tc_fence_after(); tmem_dealloc<256>(tmem); }
}

int main() {
    int* mm; float* d; uint32_t* raw;
    cudaMalloc(&mm, 4); cudaMalloc(&d, sizeof(float) * 2 * 128 * 64); cudaMalloc(&raw, 4 * 128 * 8);
    const int smem = 24576 + 8192 + 256 + 1024;
    cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    for (int rowoff : {0, 1, 11, 22})
        for (int k4 : {0, 3}) {
            cudaMemset(mm, 0, 4);
            probe<<<1, 128, smem>>>(rowoff, k4, mm, d, raw);
            cudaError_t e = cudaDeviceSynchronize();
            if (e != cudaSuccess) { printf("rowoff %d k4 %d: %s\n", rowoff, k4, cudaGetErrorString(e)); return 1; }
            int h; cudaMemcpy(&h, mm, 4, cudaMemcpyDeviceToHost);
            static float hd[2 * 128 * 64]; cudaMemcpy(hd, d, sizeof(hd), cudaMemcpyDeviceToHost);
            static uint32_t hr[128 * 8]; cudaMemcpy(hr, raw, sizeof(hr), cudaMemcpyDeviceToHost);
            int dbad = 0; double mx = 0;
            for (int i = 0; i < 128 * 64; ++i) { if (hd[i] != hd[128 * 64 + i]) ++dbad; if (fabs(hd[128 * 64 + i]) > mx) mx = fabs(hd[128 * 64 + i]); }
            printf("rowoff %2d k4 %d: cp mismatches %4d / 2048 ; D_ts != D_ss at %5d / 8192 (max |D_ss| %.0f) ; lane0 raw %08x %08x lane9 raw %08x\n", rowoff, k4, h, dbad, mx, hr[0], hr[1], hr[72]);
        }
    return 0;
}
