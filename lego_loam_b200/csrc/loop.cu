// loop.cu — loop-closure ICP on the device (see loop.cuh).
#include "loop.cuh"
#include <cooperative_groups.h>

namespace cg = cooperative_groups;

namespace llb {

namespace {

constexpr int ICP_THREADS = 256;
constexpr int ICP_TILE = 2048;             // target points staged per pass (32 KB of shared memory)
constexpr int ICP_NSUM = 18;               // n, p[3], q[3], qp[9], sum d2, (pad)

__device__ void svd3(const double A[9], double U[9], double s[3], double V[9])
{
    double B[9];
    for (int i = 0; i < 9; i++) { B[i] = A[i]; V[i] = (i % 4 == 0) ? 1.0 : 0.0; }
    for (int sweep = 0; sweep < 60; sweep++) {
        double off = 0.0;
        for (int p = 0; p < 2; p++)
            for (int q = p + 1; q < 3; q++) {
                double a = 0, b = 0, c = 0;
                for (int k = 0; k < 3; k++) { a += B[3 * k + p] * B[3 * k + p]; b += B[3 * k + q] * B[3 * k + q]; c += B[3 * k + p] * B[3 * k + q]; }
                off = fmax(off, fabs(c) / (sqrt(a * b) + DBL_MIN));
                if (fabs(c) <= 1e-300) continue;
                const double zeta = (b - a) / (2.0 * c);
                const double t = (zeta >= 0 ? 1.0 : -1.0) / (fabs(zeta) + sqrt(1.0 + zeta * zeta));
                const double cs = 1.0 / sqrt(1.0 + t * t), sn = cs * t;
                for (int k = 0; k < 3; k++) {
                    const double bp = B[3 * k + p], bq = B[3 * k + q];
                    B[3 * k + p] = cs * bp - sn * bq; B[3 * k + q] = sn * bp + cs * bq;
                    const double vp = V[3 * k + p], vq = V[3 * k + q];
                    V[3 * k + p] = cs * vp - sn * vq; V[3 * k + q] = sn * vp + cs * vq;
                }
            }
        if (off < 1e-15) break;
    }
    int ord[3] = { 0, 1, 2 };
    double n[3];
    for (int j = 0; j < 3; j++) n[j] = sqrt(B[j] * B[j] + B[3 + j] * B[3 + j] + B[6 + j] * B[6 + j]);
    for (int i = 0; i < 2; i++) for (int j = i + 1; j < 3; j++) if (n[ord[j]] > n[ord[i]]) { const int t = ord[i]; ord[i] = ord[j]; ord[j] = t; }
    double Vs[9], Us[9];
    for (int j = 0; j < 3; j++) {
        const int o = ord[j];
        s[j] = n[o];
        for (int k = 0; k < 3; k++) { Vs[3 * k + j] = V[3 * k + o]; Us[3 * k + j] = n[o] > 0 ? B[3 * k + o] / n[o] : 0.0; }
    }
    if (s[2] <= 1e-12 * s[0]) {            // vanishing singular values: complete U to an orthonormal basis
        if (s[1] <= 1e-12 * s[0]) {
            double a[3] = { Us[0], Us[3], Us[6] };
            if (s[0] <= 0) { a[0] = 1; a[1] = 0; a[2] = 0; Us[0] = 1; Us[3] = 0; Us[6] = 0; }
            double e[3] = { 0, 0, 0 }; e[fabs(a[0]) < 0.9 ? 0 : 1] = 1.0;
            const double d = e[0] * a[0] + e[1] * a[1] + e[2] * a[2];
            const double u1[3] = { e[0] - d * a[0], e[1] - d * a[1], e[2] - d * a[2] };
            const double l = sqrt(u1[0] * u1[0] + u1[1] * u1[1] + u1[2] * u1[2]);
            for (int k = 0; k < 3; k++) Us[3 * k + 1] = u1[k] / l;
        }
        const double a[3] = { Us[0], Us[3], Us[6] }, b[3] = { Us[1], Us[4], Us[7] };
        Us[2] = a[1] * b[2] - a[2] * b[1]; Us[5] = a[2] * b[0] - a[0] * b[2]; Us[8] = a[0] * b[1] - a[1] * b[0];
    }
    for (int i = 0; i < 9; i++) { U[i] = Us[i]; V[i] = Vs[i]; }
}

__device__ double det3(const double M[9])
{
    return M[0] * (M[4] * M[8] - M[5] * M[7]) - M[1] * (M[3] * M[8] - M[5] * M[6]) + M[2] * (M[3] * M[7] - M[4] * M[6]);
}

// pcl::umeyama(src, dst, false) from the sums of the correspondences (transformation_estimation_svd.hpp)
__device__ void umeyama_from_sums(const double *S, float Rt[16])
{
    const double n = S[0];
    double pm[3], qm[3], sigma[9];
    for (int a = 0; a < 3; a++) { pm[a] = S[1 + a] / n; qm[a] = S[4 + a] / n; }
    for (int r = 0; r < 3; r++) for (int c = 0; c < 3; c++) sigma[3 * r + c] = S[7 + 3 * r + c] / n - qm[r] * pm[c];
    double U[9], s[3], V[9], Sg[3] = { 1, 1, 1 };
    svd3(sigma, U, s, V);
    if (det3(sigma) < 0) Sg[2] = -1;
    int rank = 0;
    for (int i = 0; i < 3; i++) if (!(fabs(s[i]) <= fabs(s[0]) * 1e-12)) rank++;
    if (rank == 2) Sg[2] = (det3(U) * det3(V) > 0) ? 1 : -1;
    double R[9];
    for (int r = 0; r < 3; r++) for (int c = 0; c < 3; c++) {
        double v = 0; for (int k = 0; k < 3; k++) v += U[3 * r + k] * Sg[k] * V[3 * c + k];
        R[3 * r + c] = v;
    }
    for (int i = 0; i < 16; i++) Rt[i] = (i % 5 == 0) ? 1.f : 0.f;
    for (int r = 0; r < 3; r++) {
        for (int c = 0; c < 3; c++) Rt[4 * r + c] = (float)R[3 * r + c];
        Rt[4 * r + 3] = (float)(qm[r] - (R[3 * r] * pm[0] + R[3 * r + 1] * pm[1] + R[3 * r + 2] * pm[2]));
    }
}

// IterativeClosestPoint::transformCloud: tr * (x, y, z, 1), float, columns added left to right (no contraction)
__device__ __forceinline__ float4 icp_transform(const float *T, const float4 p)
{
    float4 o;
    o.x = ((T[0] * p.x + T[1] * p.y) + T[2] * p.z) + T[3];
    o.y = ((T[4] * p.x + T[5] * p.y) + T[6] * p.z) + T[7];
    o.z = ((T[8] * p.x + T[9] * p.y) + T[10] * p.z) + T[11];
    o.w = p.w;
    return o;
}

// exact 1-NN of every cur[q] in tgt: work item = (block of ICP_THREADS queries, slice of the target)
__device__ void icp_search(const float4 *__restrict__ cur, int ns, const float4 *__restrict__ tgt, int nt, int slices,
                           unsigned long long *__restrict__ nn, float4 *s_tile)
{
    const int nqb = (ns + ICP_THREADS - 1) / ICP_THREADS;
    const int per = ((nt + slices - 1) / slices + ICP_TILE - 1) / ICP_TILE * ICP_TILE;
    for (int item = blockIdx.x; item < nqb * slices; item += gridDim.x) {
        const int qb = item % nqb, sl = item / nqb;
        const int q = qb * ICP_THREADS + threadIdx.x;
        float4 p = make_float4(0.f, 0.f, 0.f, 0.f);
        if (q < ns) p = cur[q];
        float best = FLT_MAX; int bi = 0x7fffffff;
        const int lo = sl * per, hi = min(lo + per, nt);
        for (int t0 = lo; t0 < hi; t0 += ICP_TILE) {
            const int tn = min(ICP_TILE, hi - t0);
            __syncthreads();
            for (int k = threadIdx.x; k < tn; k += ICP_THREADS) s_tile[k] = __ldg(&tgt[t0 + k]);
            __syncthreads();
#pragma unroll 4
            for (int k = 0; k < tn; k++) {
                const float4 c = s_tile[k];                  // broadcast read
                const float dx = p.x - c.x, dy = p.y - c.y, dz = p.z - c.z;
                float d = dx * dx; d += dy * dy; d += dz * dz;   // flann::L2_Simple
                if (d < best) { best = d; bi = t0 + k; }     // ascending index: the first of equal distances stays
            }
        }
        if (q < ns && bi != 0x7fffffff)
            atomicMin(&nn[q], ((unsigned long long)__float_as_uint(best) << 32) | (unsigned)bi);
    }
}

__global__ void __launch_bounds__(ICP_THREADS)
icp_kernel(IcpParams prm, const float4 *__restrict__ src, int ns, const float4 *__restrict__ tgt, int nt, float4 *cur,
           unsigned long long *nn, double *partials, IcpState *st, int slices)
{
    cg::grid_group grid = cg::this_grid();
    __shared__ float4 s_tile[ICP_TILE];
    __shared__ double s_red[ICP_THREADS / 32][ICP_NSUM];
    const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
    const int gtid = blockIdx.x * ICP_THREADS + tid, gsz = gridDim.x * ICP_THREADS;
    const double max_d2 = prm.max_corr_dist * prm.max_corr_dist;

    for (int q = gtid; q < ns; q += gsz) { cur[q] = src[q]; nn[q] = ~0ull; }     // *input_transformed = *input_
    if (gtid == 0) {
        for (int i = 0; i < 16; i++) { st->T[i] = (i % 5 == 0) ? 1.f : 0.f; st->Tr[i] = st->T[i]; }
        st->prev_mse = DBL_MAX; st->fitness = DBL_MAX; st->converged = 0; st->iterations = 0; st->state = 0; st->n_corr = 0;
        st->done = (ns <= 0 || nt <= 0) ? 1 : 0;
    }
    grid.sync();
    if (ns <= 0 || nt <= 0) return;

    for (int final_pass = 0; final_pass < 2; final_pass++) {
        for (;;) {
            icp_search(cur, ns, tgt, nt, slices, nn, s_tile);
            grid.sync();
            // ---- sums over the correspondences (fitness pass: every point counts, registration.hpp getFitnessScore)
            double a[ICP_NSUM];
#pragma unroll
            for (int k = 0; k < ICP_NSUM; k++) a[k] = 0.0;
            for (int q = gtid; q < ns; q += gsz) {
                const unsigned long long key = nn[q];
                if (key == ~0ull) continue;                                      // no neighbour was found (a non-finite point)
                const float d2 = __uint_as_float((unsigned)(key >> 32));
                if (!final_pass && (double)d2 > max_d2) continue;            // correspondence_estimation.hpp
                const float4 p = cur[q], c = __ldg(&tgt[(unsigned)(key & 0xffffffffu)]);
                const double px = p.x, py = p.y, pz = p.z, qx = c.x, qy = c.y, qz = c.z;
                a[0] += 1.0; a[1] += px; a[2] += py; a[3] += pz; a[4] += qx; a[5] += qy; a[6] += qz;
                a[7] += qx * px; a[8] += qx * py; a[9] += qx * pz; a[10] += qy * px; a[11] += qy * py; a[12] += qy * pz;
                a[13] += qz * px; a[14] += qz * py; a[15] += qz * pz; a[16] += (double)d2;
            }
#pragma unroll
            for (int k = 0; k < ICP_NSUM - 1; k++) {
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) a[k] += __shfl_xor_sync(FULL, a[k], o);
                if (lane == 0) s_red[w][k] = a[k];
            }
            __syncthreads();
            if (tid < ICP_NSUM - 1) {
                double s = 0.0;
                for (int k = 0; k < ICP_THREADS / 32; k++) s += s_red[k][tid];
                partials[(size_t)blockIdx.x * ICP_NSUM + tid] = s;
            }
            grid.sync();
            if (blockIdx.x == 0) {
                if (tid < ICP_NSUM - 1) {
                    double s = 0.0;
                    for (int b = 0; b < (int)gridDim.x; b++) s += __ldcg(&partials[(size_t)b * ICP_NSUM + tid]);
                    st->sums[tid] = s;
                }
                __syncthreads();
                if (tid == 0) {
                    const double cnt = st->sums[0];
                    if (final_pass) {
                        st->fitness = cnt > 0 ? st->sums[16] / cnt : DBL_MAX;
                        st->done = 1;
                    } else if (cnt < 3.0) {                                  // min_number_correspondences_: not converged
                        st->converged = 0; st->state = 0; st->done = 1;
                        for (int i = 0; i < 16; i++) st->Tr[i] = (i % 5 == 0) ? 1.f : 0.f;
                    } else {
                        float Tr[16], Tn[16];
                        umeyama_from_sums(st->sums, Tr);
                        for (int r = 0; r < 4; r++) for (int c = 0; c < 4; c++)   // final_transformation_ = transformation_ * final
                            Tn[4 * r + c] = ((Tr[4 * r] * st->T[c] + Tr[4 * r + 1] * st->T[4 + c]) + Tr[4 * r + 2] * st->T[8 + c]) + Tr[4 * r + 3] * st->T[12 + c];
                        for (int i = 0; i < 16; i++) { st->T[i] = Tn[i]; st->Tr[i] = Tr[i]; }
                        const int it = ++st->iterations;
                        st->n_corr = (int)cnt;
                        const double mse = st->sums[16] / cnt;
                        // DefaultConvergenceCriteria::hasConverged
                        const double cos_angle = 0.5 * ((double)Tr[0] + (double)Tr[5] + (double)Tr[10] - 1.0);
                        const double tsq = (double)Tr[3] * Tr[3] + (double)Tr[7] * Tr[7] + (double)Tr[11] * Tr[11];
                        int state = 0;
                        if (it >= prm.max_iterations) state = 1;
                        else if (cos_angle >= 1.0 - prm.transformation_epsilon && tsq <= prm.transformation_epsilon) state = 2;
                        else if (fabs(mse - st->prev_mse) < 1e-12) state = 3;
                        else if (fabs(mse - st->prev_mse) / st->prev_mse < prm.fitness_epsilon) state = 4;
                        else st->prev_mse = mse;
                        if (state) { st->converged = 1; st->state = state; st->done = 1; }
                    }
                    __threadfence();
                }
            }
            grid.sync();
            const int done = __ldcg(&st->done);
            if (final_pass) break;
            // transformCloud(*input_transformed, *input_transformed, transformation_)
            float Tr[12];
#pragma unroll
            for (int i = 0; i < 12; i++) Tr[i] = __ldcg(&st->Tr[i]);
            for (int q = gtid; q < ns; q += gsz) { cur[q] = icp_transform(Tr, cur[q]); nn[q] = ~0ull; }
            grid.sync();
            if (done) break;
        }
        if (final_pass) break;
        // getFitnessScore: the ORIGINAL source through final_transformation_
        float Tf[12];
#pragma unroll
        for (int i = 0; i < 12; i++) Tf[i] = __ldcg(&st->T[i]);
        for (int q = gtid; q < ns; q += gsz) { cur[q] = icp_transform(Tf, src[q]); nn[q] = ~0ull; }
        if (gtid == 0) st->done = 0;
        grid.sync();
    }
}

__global__ void __launch_bounds__(1024)
loop_filter_intensity_kernel(const float4 *__restrict__ in, int n, float4 *__restrict__ out, int *__restrict__ n_out)
{
    __shared__ int s_scan[33];
    int base = 0;
    for (int i0 = 0; i0 < n; i0 += 1024) {
        const int i = i0 + threadIdx.x;
        float4 p = make_float4(0.f, 0.f, 0.f, 0.f);
        int keep = 0;
        if (i < n) { p = __ldg(&in[i]); keep = ((int)p.w >= 0) ? 1 : 0; }     // MO:846
        int total;
        const int pos = block_excl_scan(keep, s_scan, total);
        if (keep) out[base + pos] = p;
        base += total;
    }
    if (threadIdx.x == 0) *n_out = base;
}

}  // namespace

void IcpSolver::init()
{
    state_.ensure(1);
    int dev = 0, sms = 0, per_sm = 0;
    LLB_CUDA(cudaGetDevice(&dev));
    LLB_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    LLB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, icp_kernel, ICP_THREADS, 0));
    if (per_sm < 1) throw std::runtime_error("icp_kernel cannot be made resident");
    max_blocks_ = sms * std::min(per_sm, 2);
    partials_.ensure((size_t)max_blocks_ * ICP_NSUM);
}

void IcpSolver::release() { state_.release(); cur_.release(); nn_.release(); partials_.release(); }

int IcpSolver::run(const IcpParams &p, const float4 *src, int n_src, const float4 *tgt, int n_tgt, int max_iter_override, cudaStream_t s)
{
    if (max_blocks_ == 0) init();
    cur_.ensure(std::max(n_src, 1)); nn_.ensure(std::max(n_src, 1));
    IcpParams prm = p;
    if (max_iter_override > 0) prm.max_iterations = max_iter_override;
    const int nqb = std::max(1, div_up(n_src, ICP_THREADS));
    // enough (query block, target slice) items to give every resident CTA ~2, slices of at least one tile
    int slices = std::max(1, std::min(div_up(2 * max_blocks_, nqb), div_up(std::max(n_tgt, 1), ICP_TILE)));
    int grid = std::max(1, std::min(max_blocks_, nqb * slices));
    float4 *cur = cur_.p; unsigned long long *nn = nn_.p; double *part = partials_.p; IcpState *st = state_.p;
    void *args[] = { &prm, &src, &n_src, &tgt, &n_tgt, &cur, &nn, &part, &st, &slices };
    LLB_CUDA(cudaLaunchCooperativeKernel((const void *)icp_kernel, dim3(grid), dim3(ICP_THREADS), args, 0, s));
    return 1;
}

int launch_loop_filter_intensity(const float4 *in, int n, float4 *out, int *n_out_dev, cudaStream_t s)
{
    loop_filter_intensity_kernel<<<1, 1024, 0, s>>>(in, n, out, n_out_dev);
    LLB_CUDA(cudaGetLastError());
    return 1;
}

}  // namespace llb
