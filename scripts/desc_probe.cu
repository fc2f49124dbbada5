// desc_probe.cu -- hardware probe (not part of the library): how does tcgen05.mma address a K-major, 128-byte-swizzled
// A operand whose start address and 8-row-group stride are NOT multiples of 1024 bytes, and does TMA accept a tensor
// map whose strides are not ascending?  Needed to keep a whole activation halo tile resident in shared memory and
// read the nine 3x3 taps out of it with shifted descriptors.
//
//   test 1: A = 256 rows x 64 bf16 loaded by one 2D TMA box (dense 128-byte rows, SWIZZLE_128B); B = identity 64x64;
//           D[m][n] = A[row(m)][n] with row(m) = off + (m / 8) * (SBO / 128) + m % 8 -- for several (off, SBO,
//           base_offset) the result is compared with that expectation.
//   test 2: 4D tensor map over T[b][y][x][c] with dims ordered (c, x, b, y) (strides 512, 51200, 5120 bytes) and
//           box (64, 10, 2, 10): one load should give shared memory rows [(y * 2 + b) * 10 + x]; checked through the MMA.
//
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o desc_probe desc_probe.cu      run: ./desc_probe
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(2); } } while (0)

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory"); }
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
    uint32_t ok;
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
    return ok != 0;
}
__device__ __forceinline__ bool mbar_wait(uint32_t bar, uint32_t parity) {
    const long long t0 = clock64();
    while (!mbar_try_wait(bar, parity)) if (clock64() - t0 > 2000000000ll) return false;
    return true;
}
__device__ __forceinline__ uint64_t make_desc(uint32_t addr, uint32_t sbo_bytes, uint32_t base_offset) {
    uint64_t d = 0;
    d |= (uint64_t)((addr >> 4) & 0x3FFF);
    d |= (uint64_t)1 << 16;
    d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)(base_offset & 7) << 49;
    d |= (uint64_t)2 << 61;
    return d;
}

struct Args { int mode; int off_rows; int sbo; int base_offset_mode; float* out; int* status; };

// mode 0: tm_a is 2D (64, 256 rows), one box of 256 rows.  mode 1: tm_a is the 4D permuted map, one box (64,10,2,10).
__global__ void __launch_bounds__(128, 1) k_probe(const __grid_constant__ CUtensorMap tm_a, const __grid_constant__ CUtensorMap tm_b, const Args a) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    __shared__ __align__(8) uint64_t bar_load, bar_mma;
    __shared__ uint32_t tmem_base_sh;
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const uint32_t sa = base, sb = base + 65536;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) {
        mbar_init(smem_u32(&bar_load), 1);
        mbar_init(smem_u32(&bar_mma), 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_sh)), "r"(64) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = tmem_base_sh;
    bool ok = true;
    if (threadIdx.x == 0) {
        const uint32_t bl = smem_u32(&bar_load);
        const uint32_t a_bytes = a.mode == 0 ? 256 * 128 : 200 * 128;
        mbar_expect_tx(bl, a_bytes + 64 * 128);
        if (a.mode == 0)
            asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                         ::"r"(sa), "l"(&tm_a), "r"(bl), "r"(0), "r"(0) : "memory");
        else
            asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
                         ::"r"(sa), "l"(&tm_a), "r"(bl), "r"(0), "r"(0), "r"(0), "r"(0) : "memory");
        asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                     ::"r"(sb), "l"(&tm_b), "r"(bl), "r"(0), "r"(0) : "memory");
        ok = mbar_wait(bl, 0);
        if (ok) {
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const uint32_t start = sa + (uint32_t)a.off_rows * 128;
            const uint32_t bo = a.base_offset_mode ? ((start >> 7) & 7) : 0;
            // M = 128, N = 64, bf16 x bf16 -> fp32
            const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(64 >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
            for (int k = 0; k < 4; k++) {
                const uint64_t ad = make_desc(start + k * 32, (uint32_t)a.sbo, bo), bd = make_desc(sb + k * 32, 1024, 0);
                asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
                             ::"r"(tmem), "l"(ad), "l"(bd), "r"(idesc), "r"(k) : "memory");
            }
            asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar_mma)) : "memory");
            ok = mbar_wait(smem_u32(&bar_mma), 0);
        }
        if (!ok) *a.status = 1;
    }
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    if (*(volatile int*)a.status == 0) {
        const uint32_t taddr = tmem + ((uint32_t)(warp * 32) << 16);
        for (int c0 = 0; c0 < 64; c0 += 16) {
            uint32_t v[16];
            asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
                         : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
                           "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
                         : "r"(taddr + c0) : "memory");
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
            for (int j = 0; j < 16; j++) a.out[(warp * 32 + lane) * 64 + c0 + j] = __uint_as_float(v[j]);
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(64) : "memory");
}

static uint16_t f2bf(float f) { uint32_t u; memcpy(&u, &f, 4); return (uint16_t)(u >> 16); }

int main() {
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult q;
    CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q));
    EncodeTiledFn enc = (EncodeTiledFn)fn;
    const int SMEM = 65536 + 8192 + 1024;
    CK(cudaFuncSetAttribute(k_probe, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM));
    // ---- operands ---------------------------------------------------------------------------------
    std::vector<uint16_t> hA(256 * 64), hB(64 * 64, 0), hT(2 * 10 * 10 * 256);
    for (int r = 0; r < 256; r++) for (int c = 0; c < 64; c++) hA[r * 64 + c] = f2bf((float)((r * 64 + c) % 251));
    for (int i = 0; i < 64; i++) hB[i * 64 + i] = f2bf(1.0f);
    auto tval = [](int b, int y, int x, int c) { return (float)((((b * 10 + y) * 10 + x) * 64 + c) % 241); };
    for (int b = 0; b < 2; b++) for (int y = 0; y < 10; y++) for (int x = 0; x < 10; x++) for (int c = 0; c < 256; c++)
        hT[((b * 10 + y) * 10 + x) * 256 + c] = f2bf(c < 64 ? tval(b, y, x, c) : 7.0f);
    uint16_t *dA, *dB, *dT; float* dOut; int* dStatus;
    CK(cudaMalloc(&dA, hA.size() * 2)); CK(cudaMalloc(&dB, hB.size() * 2)); CK(cudaMalloc(&dT, hT.size() * 2));
    CK(cudaMalloc(&dOut, 128 * 64 * 4)); CK(cudaMalloc(&dStatus, 4));
    CK(cudaMemcpy(dA, hA.data(), hA.size() * 2, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(dB, hB.data(), hB.size() * 2, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(dT, hT.data(), hT.size() * 2, cudaMemcpyHostToDevice));
    CUtensorMap tmA, tmB, tmT;
    {
        cuuint64_t dims[2] = {64, 256}; cuuint64_t strides[1] = {128}; cuuint32_t box[2] = {64, 256}; cuuint32_t es[2] = {1, 1};
        CUresult r = enc(&tmA, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, dA, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                         CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) { printf("encode A failed %d\n", (int)r); return 2; }
        cuuint64_t dimsb[2] = {64, 64}; cuuint32_t boxb[2] = {64, 64};
        r = enc(&tmB, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, dB, dimsb, strides, boxb, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) { printf("encode B failed %d\n", (int)r); return 2; }
    }
    bool have_permuted = false;
    {
        cuuint64_t dims[4] = {256, 10, 2, 10};                       // (c, x, b, y)
        cuuint64_t strides[3] = {512, 51200, 5120};                  // x, b, y in bytes
        cuuint32_t box[4] = {64, 10, 2, 10}; cuuint32_t es[4] = {1, 1, 1, 1};
        CUresult r = enc(&tmT, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, dT, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                         CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        printf("permuted-stride tensor map (c,x,b,y): encode %s (%d)\n", r == CUDA_SUCCESS ? "OK" : "FAILED", (int)r);
        have_permuted = r == CUDA_SUCCESS;
    }
    std::vector<float> out(128 * 64);
    auto run = [&](const CUtensorMap& ta, int mode, int off, int sbo, int bom) -> int {
        Args a{mode, off, sbo, bom, dOut, dStatus};
        CK(cudaMemset(dStatus, 0, 4));
        CK(cudaMemset(dOut, 0xFF, 128 * 64 * 4));
        k_probe<<<1, 128, SMEM>>>(ta, tmB, a);
        CK(cudaDeviceSynchronize());
        int st = 0;
        CK(cudaMemcpy(&st, dStatus, 4, cudaMemcpyDeviceToHost));
        CK(cudaMemcpy(out.data(), dOut, out.size() * 4, cudaMemcpyDeviceToHost));
        return st;
    };
    // ---- test 1 ------------------------------------------------------------------------------------
    const int offs[] = {0, 1, 2, 3, 7, 8, 10, 11, 12, 21, 22};
    for (int sbo : {1024, 1280, 2048}) for (int bom : {0, 1}) for (int off : offs) {
        if (off + 15 * (sbo / 128) + 8 > 256) continue;
        int st = run(tmA, 0, off, sbo, bom);
        int bad = 0, first_bad = -1;
        for (int m = 0; m < 128; m++) {
            const int r = off + (m / 8) * (sbo / 128) + (m % 8);
            for (int n = 0; n < 64; n++) if (out[m * 64 + n] != (float)((r * 64 + n) % 251)) { bad++; if (first_bad < 0) first_bad = m; }
        }
        printf("test1 sbo=%4d base_offset=%s off=%2d : %s (mismatches %d, first bad row %d)%s\n", sbo, bom ? "(start>>7)&7" : "0", off,
               bad == 0 ? "MATCH" : "differ", bad, first_bad, st ? " [TIMEOUT]" : "");
    }
    // ---- test 2 ------------------------------------------------------------------------------------
    if (have_permuted) {
        for (int bom : {0, 1}) for (int ky = 0; ky < 3; ky++) for (int kx = 0; kx < 3; kx++) {
            int st = run(tmT, 1, ky * 20 + kx, 1280, bom);
            int bad = 0;
            for (int m = 0; m < 128; m++) {
                const int oy = m >> 4, b = (m >> 3) & 1, ox = m & 7;
                for (int n = 0; n < 64; n++) if (out[m * 64 + n] != tval(b, oy + ky, ox + kx, n)) bad++;
            }
            printf("test2 tap(ky=%d,kx=%d) base_offset=%s : %s (mismatches %d)%s\n", ky, kx, bom ? "(start>>7)&7" : "0", bad == 0 ? "MATCH" : "differ", bad, st ? " [TIMEOUT]" : "");
        }
    }
    return 0;
}
