// Connected-component statistics of label maps on the device (the quality-control gates that decide which
// subjects proceed: common/cardiac_utils.py:77-136 sa_pass_quality_control, :137-166 la_pass_quality_control, :1616-1652
// atrium_pass_quality_control; common/image_utils.py:227-249 get_largest_cc / remove_small_cc).
//
// One CTA per (slice, class): the slice's binary mask (label == class) is labelled in SHARED memory by min-label propagation with
// pointer jumping (every pixel starts as its own root = its linear index, which fits 16 bits for slices up to 65535 pixels; a pixel
// adopts the smallest root among its 4 / 8 neighbours; roots are chased to their fixed points, so labels only ever decrease and
// the sweep count is the longest "staircase" of a component, not its diameter), then component areas are accumulated in a second
// 16-bit array.  The kernel emits per (slice, class): area, number of components, number of components larger than `thres`
// pixels, the largest component's area and first pixel, and the area that remove_small_cc(thres) keeps -- everything the gates
// need except the union mask of one mid-cavity slice, which the host builds.  Connectivity 1 = faces (scipy.ndimage.label's
// default structure, used by get_largest_cc / remove_small_cc), 2 = faces + corners in the slice plane (skimage.measure.label(...,
// connectivity=2) on an (X, Y, 1) array, used by atrium_pass_quality_control).
#include "common.cuh"

namespace ukbb {

struct CcOut {              // matches the int32[6] rows of ukbb_cc_stats
    int area, n_cc, n_big, max_area, max_root, kept_area;
};

__global__ void __launch_bounds__(1024, 1)
cc_stats_kernel(const uint8_t* __restrict__ labels, int npix, int w, int h, const int* __restrict__ classes, int n_cls, int conn, int thres,
                int* __restrict__ out) {
    extern __shared__ uint16_t sm[];
    uint16_t* L = sm;                       // root of every pixel (0xFFFF = background)
    uint16_t* A = sm + ((npix + 1) & ~1);   // area per root
    __shared__ int s_changed, s_area, s_ncc, s_nbig, s_max, s_root, s_kept;
    const int slice = blockIdx.x / n_cls, cls = classes[blockIdx.x % n_cls];
    const uint8_t* img = labels + (size_t)slice * npix;
    if (threadIdx.x == 0) { s_area = 0; s_ncc = 0; s_nbig = 0; s_max = 0; s_root = -1; s_kept = 0; }
    for (int p = threadIdx.x; p < npix; p += blockDim.x) { L[p] = img[p] == cls ? (uint16_t)p : (uint16_t)0xFFFF; A[p] = 0; }
    __syncthreads();
    for (int sweep = 0; sweep < 4096; ++sweep) {
        if (threadIdx.x == 0) s_changed = 0;
        __syncthreads();
        bool changed = false;
        for (int p = threadIdx.x; p < npix; p += blockDim.x) {
            uint32_t m = L[p];
            if (m == 0xFFFF) continue;
            const int x = p % w, y = p / w;
#pragma unroll
            for (int dy = -1; dy <= 1; ++dy)
#pragma unroll
                for (int dx = -1; dx <= 1; ++dx) {
                    if ((dx | dy) == 0 || (conn == 1 && dx != 0 && dy != 0)) continue;
                    const int xx = x + dx, yy = y + dy;
                    if (xx < 0 || xx >= w || yy < 0 || yy >= h) continue;
                    const uint32_t q = L[yy * w + xx];
                    if (q < m) m = q;
                }
            uint32_t r = m;
            while (L[r] < r) r = L[r];                    // chase to the current root (labels only decrease: benign races)
            if (r < L[p]) { L[p] = (uint16_t)r; changed = true; }
        }
        if (changed) s_changed = 1;
        __syncthreads();
        if (!s_changed) break;
        __syncthreads();
    }
    // flatten and count
    for (int p = threadIdx.x; p < npix; p += blockDim.x) {
        uint32_t r = L[p];
        if (r == 0xFFFF) continue;
        while (L[r] < r) r = L[r];
        // two 16-bit counters share a 32-bit word: add to the half that holds A[r] (areas < 65536 by construction)
        atomicAdd(reinterpret_cast<unsigned int*>(A) + (r >> 1), (r & 1) ? 0x10000u : 1u);
        atomicAdd(&s_area, 1);
    }
    __syncthreads();
    for (int p = threadIdx.x; p < npix; p += blockDim.x) {
        const int a = A[p];
        if (a == 0) continue;
        atomicAdd(&s_ncc, 1);
        if (a > thres) atomicAdd(&s_nbig, 1);
        if (a >= thres) atomicAdd(&s_kept, a);            // remove_small_cc drops components with area < thres
        atomicMax(&s_max, a);
    }
    __syncthreads();
    for (int p = threadIdx.x; p < npix; p += blockDim.x)   // first (lowest-index) root among the largest components = scipy's lowest label
        if (A[p] == s_max && s_max > 0) atomicMin(reinterpret_cast<unsigned int*>(&s_root), (unsigned int)p);
    __syncthreads();
    if (threadIdx.x == 0) {
        int* o = out + (size_t)blockIdx.x * 6;
        o[0] = s_area; o[1] = s_ncc; o[2] = s_nbig; o[3] = s_max; o[4] = s_root; o[5] = s_kept;
    }
}

}  // namespace ukbb

extern "C" int ukbb_cc_stats(const uint8_t* labels, int n_slices, int x, int y, const int* classes_host, int n_classes, int connectivity,
                             int thres, int* stats, void* stream) {
    using namespace ukbb;
    UKBB_REQUIRE(labels && classes_host && stats, "cc_stats: null argument");
    UKBB_REQUIRE(n_slices > 0 && x > 0 && y > 0 && (long long)x * y < 65535, "cc_stats: slices of %d x %d pixels (max 65534 pixels per slice)", x, y);
    UKBB_REQUIRE(n_classes > 0 && n_classes <= 16, "cc_stats: %d classes (1..16)", n_classes);
    UKBB_REQUIRE(connectivity == 1 || connectivity == 2, "cc_stats: connectivity must be 1 (faces) or 2 (faces + corners)");
    cudaStream_t st = (cudaStream_t)stream;
    const int npix = x * y;
    const size_t smem = (size_t)2 * ((npix + 1) & ~1) * sizeof(uint16_t);
    static int* d_cls = nullptr;                                      // 16 class ids: a tiny per-process scratch
    if (!d_cls) UKBB_CUDA(cudaMalloc(&d_cls, 16 * sizeof(int)));
    UKBB_CUDA(cudaMemcpyAsync(d_cls, classes_host, n_classes * sizeof(int), cudaMemcpyHostToDevice, st));
    UKBB_CUDA(cudaFuncSetAttribute(cc_stats_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    cc_stats_kernel<<<n_slices * n_classes, 1024, smem, st>>>(labels, npix, x, y, d_cls, n_classes, connectivity, thres, stats);
    UKBB_CUDA(cudaGetLastError());
    return UKBB_OK;
}
