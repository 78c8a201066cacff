// Host-side context of one sequence: device buffers, stream, tensor maps and the launch glue.
#pragma once
#include "track.cuh"
#include <vector>

namespace lvtb
{

int launch_border_filter(const float2 *src_xy, const float *src_resp, const int *src_n, int src_stride,
                         const FeatDev *d_feats, int n_images, int rows, int cols, int *error, cudaStream_t stream);
int launch_depth_gate(const FeatDev &f, const float *d_depth, const lvt_params_c &p, cudaStream_t stream);
int track_configure(int owner_cap, TrackLaunchCfg *cfg);
int launch_track_frame(TrackState *st, void *ctl, FrameResult *result, const PointStore &map, const PointStore &staged,
                       const FeatDev *d_feats, const TrackParams &tp, const TrackScratch &sc, const CandLists &row_cand,
                       const TrackLaunchCfg &cfg, int *d_error, cudaStream_t stream, cudaEvent_t right_ready = nullptr,
                       int parts = 3, EarlyResult *early = nullptr, int early_seq = 0, const TrackOverlap *ov = nullptr,
                       const void *ctl_prev = nullptr /* FrameCtl of the frame launched before this one */);
int launch_clear_halt(TrackState *st, cudaStream_t stream);
size_t frame_ctl_bytes();
int launch_rowcand(const FeatDev *d_feats, const CamParams &cam, const CandLists &L, cudaStream_t stream);
int launch_reset_state(TrackState *st, cudaStream_t stream);
int launch_match_seam(const double *d_xyz, const uint32_t *d_pdesc, int m, const PoseD &pose, const FeatDev *d_feat,
                      const CamParams &cam, int retry_below, const MatchScratch &ms, const CandLists &lists,
                      int *d_match_idx, float *d_d1, float *d_d2, int *d_count_retried, const TrackLaunchCfg &cfg,
                      cudaStream_t stream);
int launch_row_seam(const FeatDev *d_feats, const CamParams &cam, const CandLists &lists, int *d_choice, int *d_items,
                    int *d_query, int *d_train, int *d_count, const TrackLaunchCfg &cfg, cudaStream_t stream);
int launch_pose_seam(const double *d_xyz, const float2 *d_uv, int m, const PoseD &init, const CamParams &cam,
                     uint8_t *d_level, uint8_t *d_inlier, double *d_e2, PoseD *d_out, int *d_n_inliers, cudaStream_t stream);
int launch_tri_seam(const PoseD &pose, const CamParams &cam, const float2 *d_uvl, const float2 *d_uvr, int n,
                    double *d_xyz, uint8_t *d_ok, cudaStream_t stream);

// owns every cudaMalloc of a context; freed together
struct DeviceArena
{
    std::vector<void *> blocks;
    size_t total = 0;
    template <class T>
    int alloc(T **p, size_t count)
    {
        void *q = nullptr;
        const size_t bytes = (count ? count : 1) * sizeof(T);
        if (cudaMalloc(&q, bytes) != cudaSuccess)
        {
            set_last_error(__FILE__, __LINE__, "cudaMalloc failed");
            return LVTK_ERR_CUDA;
        }
        cudaMemset(q, 0, bytes);
        blocks.push_back(q);
        total += bytes;
        *p = static_cast<T *>(q);
        return LVTK_OK;
    }
    // move an allocation to a new size (contents copied, the rest zeroed); everything on the device is idle
    template <class T>
    int regrow(T **p, size_t old_count, size_t new_count)
    {
        T *old = *p, *q = nullptr;
        if (int rc = alloc(&q, new_count))
            return rc;
        if (old && old_count)
            if (cudaMemcpy(q, old, sizeof(T) * (old_count < new_count ? old_count : new_count), cudaMemcpyDeviceToDevice) != cudaSuccess)
            {
                set_last_error(__FILE__, __LINE__, "cudaMemcpy failed");
                return LVTK_ERR_CUDA;
            }
        for (size_t i = 0; i < blocks.size(); i++)
            if (blocks[i] == old)
            {
                blocks.erase(blocks.begin() + i);
                break;
            }
        cudaFree(old);
        *p = q;
        return LVTK_OK;
    }
    void release()
    {
        for (void *q : blocks)
            cudaFree(q);
        blocks.clear();
    }
};

int make_image_pool(ImagePool *pool, DeviceArena &arena, int rows, int cols, int n_slots);
int make_detect_workspace(DetectWorkspace *ws, DeviceArena &arena, const TileGrid &grid, int rows, int pitch, int batch);
int make_feat(FeatDev *f, DeviceArena &arena, int cap, int n_cells, int rows);
int make_points(PointStore *p, DeviceArena &arena, int cap);
CamParams make_cam_params(const lvt_params_c &p);
TileGrid make_tile_grid(int img_w, int img_h, int cell);

} // namespace lvtb
