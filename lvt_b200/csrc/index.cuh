// Feature container build-up, block-level: the 25-px hash grid of lvt_image_features_struct::init
// (lvt/src/lvt_image_features_struct.cpp:35-66, .h:82-85) as a CSR, a row CSR for the stereo band
// search (:124-137) and cleared match marks (:62).  One CTA (any size) per image; shared by
// index_kernel and by the extra CTA that brief_kernel carries for it (brief.cu).
#pragma once
#include "extract.cuh"

namespace lvtb
{

// exclusive scan of arr[0..len) in shared memory, in place; arr[len] = total
__device__ inline void block_scan_array(int *arr, int len, int *s_scan)
{
    const int per = (len + blockDim.x - 1) / blockDim.x;
    const int lo = min((int)threadIdx.x * per, len), hi = min(lo + per, len);
    int sum = 0;
    for (int i = lo; i < hi; i++)
        sum += arr[i];
    int total;
    int off = block_exclusive_scan(sum, s_scan, &total);
    for (int i = lo; i < hi; i++)
    {
        const int v = arr[i];
        arr[i] = off;
        off += v;
    }
    if (threadIdx.x == 0)
        arr[len] = total;
    __syncthreads();
}

inline __host__ __device__ int index_smem_ints(const CamParams &cam) { return cam.cells_x * cam.cells_y + 1 + cam.img_h + 2; }

// s_dyn: index_smem_ints(cam) ints of shared memory, s_scan: 34 ints.  All threads of the CTA call.
__device__ inline void block_build_index(const FeatDev &f, const CamParams &cam, int *s_dyn, int *s_scan)
{
    const int n_cells = cam.cells_x * cam.cells_y, n_rows = cam.img_h + 1; // bins 0..img_h
    int *s_cell = s_dyn, *s_row = s_dyn + n_cells + 1;
    const int n = *f.n;

    for (int i = threadIdx.x; i < n_cells + 1 + n_rows + 1; i += blockDim.x)
        s_dyn[i] = 0;
    __syncthreads();
    const float cell = (float)kHashCell;
    auto cell_of = [&](float2 p) {
        // compute_hashed_index (struct.h:82-85); positions outside the grid (possible only for
        // undistorted RGB-D keypoints, undefined in the reference) are clamped
        const int hy = min(max((int)floorf(__fdiv_rn(p.y, cell)), 0), cam.cells_y - 1);
        const int hx = min(max((int)floorf(__fdiv_rn(p.x, cell)), 0), cam.cells_x - 1);
        return hy * cam.cells_x + hx;
    };
    auto row_of = [&](float2 p) { return min(max((int)floorf(p.y), 0), cam.img_h); };
    for (int i = threadIdx.x; i < n; i += blockDim.x)
    {
        const float2 p = f.xy[i];
        atomicAdd(&s_cell[cell_of(p)], 1);
        atomicAdd(&s_row[row_of(p)], 1);
        f.matched[i] = 0;
    }
    __syncthreads();
    block_scan_array(s_cell, n_cells, s_scan);
    block_scan_array(s_row, n_rows, s_scan);
    for (int i = threadIdx.x; i <= n_cells; i += blockDim.x)
        f.cell_start[i] = s_cell[i];
    for (int i = threadIdx.x; i <= n_rows; i += blockDim.x)
        f.row_start[i] = s_row[i];
    __syncthreads();
    for (int i = threadIdx.x; i < n; i += blockDim.x)
    {
        const float2 p = f.xy[i];
        f.cell_items[atomicAdd(&s_cell[cell_of(p)], 1)] = i;
        f.row_items[atomicAdd(&s_row[row_of(p)], 1)] = i;
    }
}

} // namespace lvtb
