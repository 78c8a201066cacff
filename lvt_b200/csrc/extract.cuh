// Device-side data layout of the feature front end (detect -> NMS -> ANMS -> BRIEF -> index).
#pragma once
#include "common.cuh"

namespace lvtb
{

// Pool of pitched 8-bit images in HBM: image i starts at data + i * pitch * rows.  One 3-D
// tensor map (x, y, image) per box shape covers the whole pool, so a kernel addresses any
// resident frame by its slot index.  pitch is a multiple of 128 B (TMA needs 16).
struct ImagePool
{
    uint8_t *data = nullptr;
    int rows = 0, cols = 0, pitch = 0, n_slots = 0;
    // TMA needs the innermost box coordinate 16-byte aligned (measured: an unaligned x raises an
    // illegal-instruction fault), so boxes start at an aligned x and are wide enough for any offset
    CUtensorMap tmap_score; // box 96 x 38 x 1  (64x32 output tile, halo 3, start x0-16)
    CUtensorMap tmap_patch; // box 80 x 57 x 1  (57-px BRIEF support at byte offset 0..15)
    size_t slot_bytes() const { return (size_t)pitch * rows; }
};

// one camera of lvt_set_rectification, ready for the device: (P R)^-1, raw pinhole, distortion
struct RectifyDev
{
    double ir[9];
    double fx, fy, cx, cy, k1, k2, p1, p2, k3;
};
int make_rectify_dev(const lvt_rectify_c &r, RectifyDev *out);
// n_images (1 or 2) raw images, pitched like the pool -> their pool slots; image i is seen by *cam[i]
int launch_rectify(const uint8_t *const raw[2], uint8_t *const dst[2], int n_images, const RectifyDev *const cam[2],
                   const ImagePool &pool, cudaStream_t stream);
int launch_rectify_maps(const RectifyDev &r, int rows, int cols, float *d_map_x, float *d_map_y, cudaStream_t stream);
// tightly packed rows (cols bytes each) -> rows of `pitch` bytes (pitch a multiple of 16, >= cols)
int launch_repitch(const uint8_t *d_packed, uint8_t *d_dst, int rows, int cols, int pitch, cudaStream_t stream);

constexpr int kScoreTileW = 64, kScoreTileH = 32;
constexpr int kScoreBoxW = 96, kScoreBoxH = kScoreTileH + 6, kScoreBoxX = 16;
constexpr int kPatchW = 80, kPatchH = 57;

// detection tile grid (lvt/src/lvt_image_features_handler.cpp:95-114)
struct TileGrid
{
    int cell, nx, ny, img_w, img_h;
    __host__ __device__ int count() const { return nx * ny; }
    __host__ __device__ int tile_w(int tx) const { return (tx == nx - 1 && (tx + 1) * cell > img_w) ? img_w - tx * cell : cell; }
    __host__ __device__ int tile_h(int ty) const { return (ty == ny - 1 && (ty + 1) * cell > img_h) ? img_h - ty * cell : cell; }
};

// One image's features as the matcher consumes them.  All pointers are device memory.
struct FeatDev
{
    int *n;            // count N' after the BRIEF border filter
    float2 *xy;        // [cap]
    float *resp;       // [cap] AGAST response (0 for external corners)
    uint32_t *desc;    // [cap][8]  256-bit BRIEF
    uint8_t *matched;  // [cap] lvt_image_features_struct::m_matched_marks
    float *depth;      // [cap] RGB-D only
    int *cell_start;   // [cells+1] CSR over the 25-px hash grid (struct.cpp:58-61)
    int *cell_items;   // [cap] feature indices grouped by cell
    int *row_start;    // [rows+2] CSR over floor(y) (row matching band, struct.cpp:124-137)
    int *row_items;    // [cap]
    int cap;
};

// Scratch for `batch` images in flight through the detector.  Index [b] = image in batch.
struct DetectWorkspace
{
    int batch = 0;
    int n_tiles = 0;
    int tile_cap = 0;       // entries per tile list (power of two, >= worst-case survivors)
    uint8_t *score = nullptr;   // [batch][rows][pitch]   corner score, 0 = none
    int *parent = nullptr;      // [batch][rows][pitch]   union-find scratch of the NMS fallback
    uint32_t *tile_list = nullptr;   // [batch][n_tiles][tile_cap]  survivors, (raster << 8 | response)
    uint32_t *tile_aux = nullptr;    // [batch][n_tiles][2*tile_cap] large-tile scratch (radii, permutation)
    uint32_t *tile_out = nullptr;    // [batch][n_tiles][tile_cap]  ordered output, (y << 20 | x << 8 | response)
    int *tile_count = nullptr;       // [batch][n_tiles]
    int *tile_overflow = nullptr;    // [batch][n_tiles]  big candidates queued for tile_kernel | sequential-fallback flag
    uint32_t *big_list = nullptr;    // [batch][n_tiles][64] local maxima of components beyond the NMS kernel's buffers
    int *tile_out_count = nullptr;   // [batch][n_tiles]
    int *tiles_done = nullptr;       // [batch] tile CTAs of the image that have finished (the last one gathers)
    int *cand_count = nullptr;       // [batch] local maxima queued for nms_resolve_kernel (list lives in `parent`)
    int *retry = nullptr;            // [batch] fewer than 200 corners: redo at the lowered threshold
    int *error = nullptr;            // [1] sticky capacity error flag
};

struct DetectParams
{
    TileGrid grid;
    int threshold;     // agast_threshold
    int threshold_low; // (int)(threshold * 0.5 + 0.5)
    int max_per_cell;  // max_keypoints_per_cell
    int pitch, rows, cols;
};

// host launchers (detect.cu / brief.cu / index in match.cu)
int launch_detect(const ImagePool &pool, const DetectWorkspace &ws, const DetectParams &dp, const int *d_slots,
                  int n_images, FeatDev *d_feats /* device array [n_images] */, int border, int single_tile_nms,
                  cudaStream_t stream);
// index_cam != nullptr: one extra CTA per image builds the feature index (hash grid, row CSR, cleared
// marks) next to the descriptors, replacing a launch_index call; requires brief_can_index(cam)
int launch_brief(const ImagePool &pool, const int *d_slots, int n_images, const FeatDev *d_feats, const uint32_t *d_offsets,
                 cudaStream_t stream, const CamParams *index_cam = nullptr);
bool brief_can_index(const CamParams &cam);
int launch_index(const FeatDev *d_feats, int n_images, const CamParams &cam, cudaStream_t stream);
// pairs == nullptr: the built-in table (brief_pairs.inc); h: what brief_kernel reads (BriefArgs::offsets)
void make_brief_offsets(const signed char pairs[256][4], uint32_t h[8 * 32]);

} // namespace lvtb
