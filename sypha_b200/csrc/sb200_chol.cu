// sb200_chol.cu - dense FP64 Cholesky of the normal matrix and the triangular solves.
//
// Replaces the reference's per-iteration dense LU of the full (2n+m)^2 KKT matrix
// (/root/reference/src/sypha_solver_dense_linear.cpp:150-203: template restore + cusolverDnDgetrf,
// then cusolverDnDgetrs twice) with an m x m Cholesky of M = A D A'.
//
// Layout: row-major, lower triangle, leading dimension ld = m rounded up to 64, identity pad.
//
// Factorisation (right-looking, 64-wide panels, 2 launches per panel):
//   k_trsm_panel : X = A_ik (L_kk^-1)' as a 64x64x64 FP64 tensor-core GEMM against the pre-inverted
//                  diagonal block
//   k_update     : A_ij -= L_ik L_jk'  on 64x64 tiles with FP64 tensor-core MMA
//                  (mma.sync.m8n8k4.f64 -> SASS DMMA.8x8x4); the CTA that owns the next diagonal
//                  tile factors AND inverts it in shared memory before writing it back (look-ahead),
//                  so no separate potrf launch sits on the critical path.
// Solves (one launch each, data-flow): every 64-row block is owned by one CTA which accumulates
//   its right-hand side as the blocks it depends on are published through release/acquire flags,
//   then multiplies by the pre-inverted 64x64 diagonal block.  All CTAs are co-resident
//   (cooperative launch); spins are bounded and raise an error flag instead of hanging.
#include "sb200_kernels.cuh"
#include "sb200_chol.cuh"
#include "sb200_dmma.cuh"
#include "sb200_tile64.cuh"

namespace sb200 {

__device__ __forceinline__ void report_fail(int *info, int fail, int base)
{
    if (fail && threadIdx.x == 0)
        atomicCAS(info, 0, base + fail);
}

// write L (lower) back to the matrix and L^-1 to the inverse store
__device__ __forceinline__ void store_factored_tile(const unsigned char *smem, double *__restrict__ Atile, int ld,
                                                    double *__restrict__ linv_k, int tid)
{
    const double(*Ls)[LP] = reinterpret_cast<const double(*)[LP]>(smem + SM_LS);
    const double(*Li)[LP] = reinterpret_cast<const double(*)[LP]>(smem + SM_LI);
    for (int idx = tid; idx < TB * TB; idx += NT_TILE)
    {
        const int r = idx >> 6, c = idx & 63;
        if (c <= r) Atile[(size_t)r * ld + c] = Ls[r][c];
        linv_k[idx] = Li[r][c];
    }
}

// first diagonal tile
__global__ void __launch_bounds__(NT_TILE) k_potrf_first(double *__restrict__ A, int ld, double *__restrict__ linv,
                                                         int *info)
{
    extern __shared__ __align__(16) unsigned char dyn_smem[];
    double(*Ls)[LP] = reinterpret_cast<double(*)[LP]>(dyn_smem + SM_LS);
    const int tid = threadIdx.x;
    for (int idx = tid; idx < TB * TB; idx += NT_TILE)
    {
        const int r = idx >> 6, c = idx & 63;
        Ls[r][c] = A[(size_t)r * ld + c];
    }
    __syncthreads();
    const int fail = potrf_inv_tile64(dyn_smem, tid);
    store_factored_tile(dyn_smem, A, ld, linv, tid);
    report_fail(info, fail, 0);
}

// X = A_ik L_kk^-T = A_ik (L_kk^-1)' for every tile row i > k: one CTA per 64x64 tile, DMMA GEMM
__global__ void __launch_bounds__(128) k_trsm_panel(double *__restrict__ A, int ld, int k,
                                                    const double *__restrict__ linv)
{
    __shared__ __align__(16) double smem[2 * TB * KP];
    double(*As)[KP] = reinterpret_cast<double(*)[KP]>(smem);
    double(*Bs)[KP] = reinterpret_cast<double(*)[KP]>(smem + TB * KP);
    const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
    const int wm = w >> 1, wn = w & 1, g = lane >> 2, tg = lane & 3;
    const size_t r0 = ((size_t)k + 1 + blockIdx.x) * TB, k0 = (size_t)k * TB;
    const double *Lk = linv + (size_t)k * TB * TB;
    double acc[4][4][2];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j)
            acc[i][j][0] = acc[i][j][1] = 0.0;
#pragma unroll
    for (int kc = 0; kc < TB; kc += KC)
    {
        __syncthreads();
        load_tile_64xKC(As, A + r0 * ld + k0 + kc, ld, tid, 128, nullptr);
        load_tile_64xKC(Bs, Lk + kc, TB, tid, 128, nullptr);
        __syncthreads();
        warp_mma_32x32(As, Bs, wm, wn, lane, 1.0, acc);
    }
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j)
            *reinterpret_cast<double2 *>(A + (r0 + wm * 32 + i * 8 + g) * ld + k0 + wn * 32 + j * 8 + tg * 2) =
                make_double2(acc[i][j][0], acc[i][j][1]);
}

// trailing update with look-ahead factorisation (+ inverse) of the next diagonal tile
__global__ void __launch_bounds__(NT_TILE) k_update(double *__restrict__ A, int ld, int k, int T,
                                                    double *__restrict__ linv, int *info)
{
    extern __shared__ __align__(16) unsigned char dyn_smem[];
    double(*As)[KP] = reinterpret_cast<double(*)[KP]>(dyn_smem);
    double(*Bs)[KP] = reinterpret_cast<double(*)[KP]>(dyn_smem + TB * KP * 8);

    // decode (ti, tj), tj <= ti, both relative to k+1
    const int p = blockIdx.x;
    int ri = (int)((sqrt(8.0 * p + 1.0) - 1.0) * 0.5);
    while ((ri + 1) * (ri + 2) / 2 <= p) ++ri;
    while (ri * (ri + 1) / 2 > p) --ri;
    const int rj = p - ri * (ri + 1) / 2;
    const int ti = k + 1 + ri, tj = k + 1 + rj;
    if (ti >= T) return;

    const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
    const int row0 = (w >> 2) * 32, col0 = (w & 3) * 16;     // 8 warps: 2 x 4, warp tile 32 x 16
    const int g = lane >> 2, tg = lane & 3;
    const size_t r0 = (size_t)ti * TB, c0 = (size_t)tj * TB, k0 = (size_t)k * TB;

    const bool mma_warp = w < 8;           // warp 8 only takes part in the tile factorisation
    double acc[4][2][2];
    if (mma_warp)
    {
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < 2; ++j)
            {
                const double2 v = *reinterpret_cast<const double2 *>(
                    A + (r0 + row0 + i * 8 + g) * ld + c0 + col0 + j * 8 + tg * 2);
                acc[i][j][0] = v.x;
                acc[i][j][1] = v.y;
            }
    }
#pragma unroll
    for (int kc = 0; kc < TB; kc += KC)
    {
        __syncthreads();
        load_tile_64xKC(As, A + r0 * ld + k0 + kc, ld, tid, NT_TILE, nullptr);
        load_tile_64xKC(Bs, A + c0 * ld + k0 + kc, ld, tid, NT_TILE, nullptr);
        __syncthreads();
        if (mma_warp) warp_mma<4, 2>(As, Bs, row0, col0, lane, -1.0, acc);
    }

    if (ri == 0 && rj == 0)
    {   // next diagonal tile: factor + invert it before it goes back to memory
        __syncthreads();
        double(*Ls)[LP] = reinterpret_cast<double(*)[LP]>(dyn_smem + SM_LS);
        if (mma_warp)
        {
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 2; ++j)
                {
                    const int r = row0 + i * 8 + g, c = col0 + j * 8 + tg * 2;
                    Ls[r][c] = acc[i][j][0];
                    Ls[r][c + 1] = acc[i][j][1];
                }
        }
        __syncthreads();
        const int fail = potrf_inv_tile64(dyn_smem, tid);
        store_factored_tile(dyn_smem, A + r0 * ld + c0, ld, linv + (size_t)ti * TB * TB, tid);
        report_fail(info, fail, (int)r0);
        return;
    }
    if (mma_warp)
    {
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < 2; ++j)
                *reinterpret_cast<double2 *>(A + (r0 + row0 + i * 8 + g) * ld + c0 + col0 + j * 8 + tg * 2) =
                    make_double2(acc[i][j][0], acc[i][j][1]);
    }
}

__global__ void k_pad_identity(int n, double *__restrict__ A, int ld)
{
    // rows n..ld-1: zero, unit diagonal; columns n..ld-1 of the real rows: zero
    const long long total = (long long)ld * ld;
    for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total;
         idx += (long long)gridDim.x * blockDim.x)
    {
        const int r = (int)(idx / ld), c = (int)(idx % ld);
        if (r >= n || c >= n)
            A[idx] = (r == c) ? 1.0 : 0.0;
    }
}
void launch_pad_identity(int n, double *a, int ld, cudaStream_t st)
{
    if (ld == n) return;
    k_pad_identity<<<grid_for((long long)ld * ld, 256), 256, 0, st>>>(n, a, ld);
    ++g_launch_count;
}

// ---------------------------------------------------------------------------------------------
// data-flow triangular solves
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ int ld_acquire(const int *p)
{
    int v;
    asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release(int *p, int v)
{
    asm volatile("st.release.gpu.global.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
// thread 0 spins (bounded), then the block synchronises
__device__ __forceinline__ void wait_flag(const int *flag, int epoch, int *err)
{
    if (threadIdx.x == 0)
    {
        long long spins = 0;
        while (ld_acquire(flag) != epoch)
        {
            if (++spins > (1ll << 22))
            {
                atomicExch(err, 1);
                break;
            }
        }
    }
    __syncthreads();
}

struct TrsvCtl
{
    int *flags;        // [T]
    int *epoch;        // device counter: flags carry *epoch+1 when published in this launch
    unsigned *exits;   // CTAs that finished (last one bumps *epoch)
    int *err;
};

__device__ __forceinline__ void trsv_epilogue(const TrsvCtl &C)
{
    __syncthreads();
    if (threadIdx.x == 0)
    {
        __threadfence();
        const unsigned t = atomicAdd(C.exits, 1u);
        if (t == gridDim.x - 1)
        {
            *C.exits = 0u;
            __threadfence();
            atomicAdd(C.epoch, 1);
        }
    }
}

// forward: L y = b, in place.  256 threads: (row = t>>2, q = t&3), q splits the 64 columns.
__global__ void __launch_bounds__(256)
k_trsv_fwd(const double *__restrict__ L, int ld, const double *__restrict__ linv, double *b, int T,
           TrsvCtl C)
{
    __shared__ double rs[TB];
    const int tid = threadIdx.x, row = tid >> 2, q = tid & 3;
    const int epoch = *C.epoch + 1;
    for (int i = blockIdx.x; i < T; i += gridDim.x)
    {
        const size_t grow = (size_t)i * TB + row;
        // pre-inverted diagonal block row (lower part only matters)
        double li[16];
        {
            const double2 *src = reinterpret_cast<const double2 *>(linv + ((size_t)i * TB + row) * TB + q * 16);
#pragma unroll
            for (int j = 0; j < 8; ++j)
            {
                const double2 v = src[j];
                li[2 * j] = v.x;
                li[2 * j + 1] = v.y;
            }
        }
        double acc = 0.0;
        double cur[16];
        if (i > 0)
        {
            const double2 *src = reinterpret_cast<const double2 *>(L + grow * ld + q * 16);
#pragma unroll
            for (int j = 0; j < 8; ++j)
            {
                const double2 v = src[j];
                cur[2 * j] = v.x;
                cur[2 * j + 1] = v.y;
            }
        }
        for (int k = 0; k < i; ++k)
        {
            double nxt[16];
            if (k + 1 < i)
            {
                const double2 *src =
                    reinterpret_cast<const double2 *>(L + grow * ld + (size_t)(k + 1) * TB + q * 16);
#pragma unroll
                for (int j = 0; j < 8; ++j)
                {
                    const double2 v = src[j];
                    nxt[2 * j] = v.x;
                    nxt[2 * j + 1] = v.y;
                }
            }
            wait_flag(C.flags + k, epoch, C.err);
            const double *yk = b + (size_t)k * TB + q * 16;
#pragma unroll
            for (int j = 0; j < 16; ++j)
                acc -= cur[j] * __ldcg(yk + j);
            if (k + 1 < i)
            {
#pragma unroll
                for (int j = 0; j < 16; ++j)
                    cur[j] = nxt[j];
            }
        }
        acc += __shfl_xor_sync(0xffffffffu, acc, 1);
        acc += __shfl_xor_sync(0xffffffffu, acc, 2);
        __syncthreads();
        if (q == 0) rs[row] = __ldcg(b + grow) + acc;
        __syncthreads();
        double yv = 0.0;
#pragma unroll
        for (int j = 0; j < 16; ++j)
            yv += li[j] * rs[q * 16 + j];
        yv += __shfl_xor_sync(0xffffffffu, yv, 1);
        yv += __shfl_xor_sync(0xffffffffu, yv, 2);
        if (q == 0) b[grow] = yv;
        __threadfence();
        __syncthreads();
        if (tid == 0) st_release(C.flags + i, epoch);
    }
    trsv_epilogue(C);
}

// backward: L' x = y, in place.  256 threads: (col = t&63, q = t>>6), q splits the 64 rows.
__global__ void __launch_bounds__(256)
k_trsv_bwd(const double *__restrict__ L, int ld, const double *__restrict__ linv, double *b, int T,
           TrsvCtl C)
{
    __shared__ double xs[TB];
    __shared__ double red[4][TB];
    const int tid = threadIdx.x, col = tid & 63, q = tid >> 6;
    const int epoch = *C.epoch + 1;
    for (int ii = blockIdx.x; ii < T; ii += gridDim.x)
    {
        const int i = T - 1 - ii;
        // column `col` of Linv_ii' = row-slice of Linv_ii: rows q*16.., column col
        double li[16];
#pragma unroll
        for (int j = 0; j < 16; ++j)
            li[j] = linv[((size_t)i * TB + q * 16 + j) * TB + col];
        double acc = 0.0;
        double cur[16];
        if (i < T - 1)
        {
#pragma unroll
            for (int j = 0; j < 16; ++j)
                cur[j] = L[((size_t)(T - 1) * TB + q * 16 + j) * ld + (size_t)i * TB + col];
        }
        for (int k = T - 1; k > i; --k)
        {
            double nxt[16];
            if (k - 1 > i)
            {
#pragma unroll
                for (int j = 0; j < 16; ++j)
                    nxt[j] = L[((size_t)(k - 1) * TB + q * 16 + j) * ld + (size_t)i * TB + col];
            }
            wait_flag(C.flags + k, epoch, C.err);
            if (tid < TB) xs[tid] = __ldcg(b + (size_t)k * TB + tid);
            __syncthreads();
#pragma unroll
            for (int j = 0; j < 16; ++j)
                acc -= cur[j] * xs[q * 16 + j];
            if (k - 1 > i)
            {
#pragma unroll
                for (int j = 0; j < 16; ++j)
                    cur[j] = nxt[j];
            }
            __syncthreads();
        }
        red[q][col] = acc;
        __syncthreads();
        if (tid < TB)
            xs[tid] = __ldcg(b + (size_t)i * TB + tid) + red[0][tid] + red[1][tid] + red[2][tid] + red[3][tid];
        __syncthreads();
        double xv = 0.0;
#pragma unroll
        for (int j = 0; j < 16; ++j)
            xv += li[j] * xs[q * 16 + j];
        __syncthreads();
        red[q][col] = xv;
        __syncthreads();
        if (tid < TB)
            b[(size_t)i * TB + tid] = red[0][tid] + red[1][tid] + red[2][tid] + red[3][tid];
        __threadfence();
        __syncthreads();
        if (tid == 0) st_release(C.flags + i, epoch);
    }
    trsv_epilogue(C);
}

// ---------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------
int chol_work_ensure(ErrorSink &err, CholWork &W, int n_pad)
{
    const int T = n_pad / TB;
    if (T <= W.t_cap) return SB200_OK;
    chol_work_free(W);
    SB200_CUDA_TRY(err, cudaMalloc(&W.linv, sizeof(double) * (size_t)T * TB * TB));
    SB200_CUDA_TRY(err, cudaMalloc(&W.ctl, sizeof(int) * (size_t)(2 * T + 16)));
    SB200_CUDA_TRY(err, cudaMemset(W.ctl, 0, sizeof(int) * (size_t)(2 * T + 16)));
    W.t_cap = T;
    int dev = 0, sms = 0, occ_f = 0, occ_b = 0;
    SB200_CUDA_TRY(err, cudaGetDevice(&dev));
    SB200_CUDA_TRY(err, cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    SB200_CUDA_TRY(err, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ_f, k_trsv_fwd, 256, 0));
    SB200_CUDA_TRY(err, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ_b, k_trsv_bwd, 256, 0));
    const int occ = occ_f < occ_b ? occ_f : occ_b;
    W.max_coop_grid = sms * (occ < 1 ? 1 : occ);
    SB200_CUDA_TRY(err, cudaFuncSetAttribute(k_potrf_first, cudaFuncAttributeMaxDynamicSharedMemorySize, SM_TOTAL));
    SB200_CUDA_TRY(err, cudaFuncSetAttribute(k_update, cudaFuncAttributeMaxDynamicSharedMemorySize, SM_TOTAL));
    return SB200_OK;
}
void chol_work_free(CholWork &W)
{
    if (W.linv) cudaFree(W.linv);
    if (W.ctl) cudaFree(W.ctl);
    W = CholWork{};
}

void launch_potrf(CholWork &W, int n, double *a, int ld, int *info, cudaStream_t st)
{
    (void)n;
    const int T = ld / TB;
    k_potrf_first<<<1, NT_TILE, SM_TOTAL, st>>>(a, ld, W.linv, info);
    ++g_launch_count;
    for (int k = 0; k + 1 < T; ++k)
    {
        const int rem = T - 1 - k;
        k_trsm_panel<<<rem, 128, 0, st>>>(a, ld, k, W.linv);
        k_update<<<rem * (rem + 1) / 2, NT_TILE, SM_TOTAL, st>>>(a, ld, k, T, W.linv, info);
        g_launch_count += 2;
    }
}

static cudaError_t launch_coop(const void *fn, int grid, int block, void **args, cudaStream_t st)
{
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(grid);
    cfg.blockDim = dim3(block);
    cfg.dynamicSmemBytes = 0;
    cfg.stream = st;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeCooperative;
    at[0].val.cooperative = 1;
    cfg.attrs = at;
    cfg.numAttrs = 1;
    return cudaLaunchKernelExC(&cfg, fn, args);
}

void launch_potrs(CholWork &W, int n, const double *l, int ld, double *b, cudaStream_t st)
{
    (void)n;
    int T = ld / TB;
    const int grid = T < W.max_coop_grid ? T : W.max_coop_grid;
    int *ctl = W.ctl;
    TrsvCtl Cf{ctl + 16, ctl + 0, reinterpret_cast<unsigned *>(ctl + 1), ctl + 4};
    TrsvCtl Cb{ctl + 16 + W.t_cap, ctl + 2, reinterpret_cast<unsigned *>(ctl + 3), ctl + 4};
    const double *linv = W.linv;
    {
        void *args[] = {(void *)&l, (void *)&ld, (void *)&linv, (void *)&b, (void *)&T, (void *)&Cf};
        launch_coop((const void *)k_trsv_fwd, grid, 256, args, st);
    }
    {
        void *args[] = {(void *)&l, (void *)&ld, (void *)&linv, (void *)&b, (void *)&T, (void *)&Cb};
        launch_coop((const void *)k_trsv_bwd, grid, 256, args, st);
    }
    g_launch_count += 2;
}

} // namespace sb200
