// Host side of liblvt_b200.so: the per-sequence context, the seam ABI (include/lvt_kernels.h),
// the lvt_system mirror and the public C ABI (include/lvt_c.h).
//
// lvt_system's control flow (lvt/src/lvt_system.cpp:157-207) is kept on the host only as far as
// the C ABI needs it (LOST short-circuit, frame counter, returning the pose); everything
// perform_tracking does runs on the device (track.cu).  There is no CPU implementation of any
// stage in this library: without a CUDA device lvt_create returns NULL and the seam calls
// return LVTK_ERR_NO_DEVICE.
#define LVT_EXPORT_FUNCTIONS
#include "context.cuh"
#include "upload.cuh"
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <mutex>
#include <string>

namespace lvtb
{

static thread_local char g_err[512] = "";
void set_last_error(const char *file, int line, const char *what)
{
    std::snprintf(g_err, sizeof(g_err), "%s:%d: %s", file, line, what);
    if (std::getenv("LVT_B200_VERBOSE"))
        std::fprintf(stderr, "[lvt_b200] %s\n", g_err);
}
const char *last_error() { return g_err; }
// ---- per-kernel device time: CUDA event pairs recorded on the launching stream ---------------
namespace
{
struct Profiler
{
    std::atomic<bool> on{false};
    std::mutex mu; // held from prof_begin to prof_end: with profiling on, launches of all threads serialise
    std::vector<cudaEvent_t> pool;
    std::vector<int> ids; // kernel id of pair i (events 2i, 2i+1)
    size_t used = 0;
    double ms[K_COUNT] = {};
    long count[K_COUNT] = {};
} g_prof;
} // namespace
static std::atomic<long> g_launches{0};
void count_launch() { g_launches.fetch_add(1, std::memory_order_relaxed); }
long launch_count() { return g_launches.load(std::memory_order_relaxed); }
bool prof_enabled() { return g_prof.on.load(std::memory_order_relaxed); }
void prof_begin(cudaStream_t s, int id)
{
    g_prof.mu.lock();
    if (g_prof.pool.size() < 2 * (g_prof.used + 1))
    {
        cudaEvent_t a, b;
        cudaEventCreate(&a);
        cudaEventCreate(&b);
        g_prof.pool.push_back(a);
        g_prof.pool.push_back(b);
        g_prof.ids.push_back(id);
    }
    g_prof.ids[g_prof.used] = id;
    cudaEventRecord(g_prof.pool[2 * g_prof.used], s);
}
void prof_end(cudaStream_t s, int)
{
    cudaEventRecord(g_prof.pool[2 * g_prof.used + 1], s);
    g_prof.used++;
    g_prof.mu.unlock();
}
static void prof_collect()
{
    std::lock_guard<std::mutex> lk(g_prof.mu);
    for (size_t i = 0; i < g_prof.used; i++)
    {
        float t = 0;
        if (cudaEventElapsedTime(&t, g_prof.pool[2 * i], g_prof.pool[2 * i + 1]) == cudaSuccess)
        {
            g_prof.ms[g_prof.ids[i]] += t;
            g_prof.count[g_prof.ids[i]]++;
        }
    }
    g_prof.used = 0;
}
bool debug_sync_enabled()
{
    static const bool on = std::getenv("LVT_B200_SYNC") != nullptr;
    return on;
}

// ---------------------------------------------------------------------------------------------
// parameters (lvt/src/lvt_parameters.cpp:29-93)
// ---------------------------------------------------------------------------------------------
static void params_default(lvt_params_c *p)
{
    std::memset(p, 0, sizeof(*p));
    p->fx = p->fy = p->cx = p->cy = 0.5f;
    p->near_plane_distance = 0.1f;
    p->far_plane_distance = 500.0f;
    p->triangulation_ratio_test_threshold = 0.60f;
    p->tracking_ratio_test_threshold = 0.80f;
    p->descriptor_matching_threshold = 30.0f;
    p->min_num_matches_for_tracking = 10;
    p->tracking_radius = 25;
    p->agast_threshold = 25;
    p->untracked_threshold = 10;
    p->staged_threshold = 2;
    p->detection_cell_size = 250;
    p->max_keypoints_per_cell = 150;
    p->triangulation_policy = 1;
    p->enable_logging = 1;
    p->enable_visualization = 0;
    p->viewer_camera_size = 0.6f;
    p->viewer_point_size = 5;
}

// flat "key: value" reader for OpenCV FileStorage YAML 1.0; absent keys read as 0 like cv::FileNode
static int params_from_file(lvt_params_c *p, const char *file)
{
    FILE *f = file ? std::fopen(file, "r") : nullptr;
    if (!f)
        return 0;
    std::map<std::string, double> kv;
    char line[1024];
    while (std::fgets(line, sizeof(line), f))
    {
        char *hash = std::strchr(line, '#');
        if (hash)
            *hash = 0;
        if (line[0] == '%' || line[0] == '-')
            continue;
        char *colon = std::strchr(line, ':');
        if (!colon)
            continue;
        *colon = 0;
        std::string key(line);
        const size_t b = key.find_first_not_of(" \t"), e = key.find_last_not_of(" \t\r\n");
        if (b == std::string::npos)
            continue;
        key = key.substr(b, e - b + 1);
        char *end = nullptr;
        const double v = std::strtod(colon + 1, &end);
        if (end == colon + 1)
            continue;
        kv[key] = v;
    }
    std::fclose(f);
    auto num = [&kv](const char *k) {
        auto it = kv.find(k);
        return it == kv.end() ? 0.0 : it->second;
    };
    auto integer = [&num](const char *k) { return (int)std::lrint(num(k)); };
    std::memset(p, 0, sizeof(*p));
    p->fx = (float)num("fx");
    p->fy = (float)num("fy");
    p->cx = (float)num("cx");
    p->cy = (float)num("cy");
    p->k1 = (float)num("k1");
    p->k2 = (float)num("k2");
    p->p1 = (float)num("p1");
    p->p2 = (float)num("p2");
    p->k3 = (float)num("k3");
    p->baseline = (float)num("baseline");
    p->img_width = integer("img_width");
    p->img_height = integer("img_height");
    p->near_plane_distance = (float)num("near_plane_distance");
    p->far_plane_distance = (float)num("far_plane_distance");
    p->triangulation_ratio_test_threshold = (float)num("triangulation_ratio_test_threshold");
    p->tracking_ratio_test_threshold = (float)num("tracking_ratio_test_threshold");
    p->min_num_matches_for_tracking = integer("min_num_matches_for_tracking");
    p->tracking_radius = integer("tracking_radius");
    p->agast_threshold = integer("agast_threshold");
    p->untracked_threshold = integer("untracked_threshold");
    p->staged_threshold = integer("staged_threshold");
    p->descriptor_matching_threshold = (float)num("descriptor_matching_threshold");
    p->detection_cell_size = integer("detection_cell_size");
    p->max_keypoints_per_cell = integer("max_keypoints_per_cell");
    p->enable_logging = integer("enable_logging") != 0;
    p->enable_visualization = integer("enable_visualization") != 0;
    p->triangulation_policy = integer("triangulation_policy");
    p->viewer_camera_size = (float)num("viewer_camera_size");
    p->viewer_point_size = integer("viewer_point_size");
    return 1;
}

// ---------------------------------------------------------------------------------------------
// buffers
// ---------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                  const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static int encode_pool_map(CUtensorMap *map, const ImagePool &pool, int box_w, int box_h)
{
    static EncodeTiledFn encode = nullptr;
    if (!encode)
    {
        void *fn = nullptr;
        cudaDriverEntryPointQueryResult qres;
        LVT_CUDA_TRY(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres));
        if (!fn || qres != cudaDriverEntryPointSuccess)
        {
            set_last_error(__FILE__, __LINE__, "cuTensorMapEncodeTiled not available");
            return LVTK_ERR_CUDA;
        }
        encode = reinterpret_cast<EncodeTiledFn>(fn);
    }
    const cuuint64_t gdim[3] = {(cuuint64_t)pool.cols, (cuuint64_t)pool.rows, (cuuint64_t)pool.n_slots};
    const cuuint64_t gstride[2] = {(cuuint64_t)pool.pitch, (cuuint64_t)pool.pitch * pool.rows};
    const cuuint32_t box[3] = {(cuuint32_t)box_w, (cuuint32_t)box_h, 1};
    const cuuint32_t estride[3] = {1, 1, 1};
    const CUresult r = encode(map, CU_TENSOR_MAP_DATA_TYPE_UINT8, 3, pool.data, gdim, gstride, box, estride,
                              CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                              CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS)
    {
        set_last_error(__FILE__, __LINE__, "cuTensorMapEncodeTiled failed");
        return LVTK_ERR_CUDA;
    }
    return LVTK_OK;
}

int make_image_pool(ImagePool *pool, DeviceArena &arena, int rows, int cols, int n_slots)
{
    pool->rows = rows;
    pool->cols = cols;
    pool->pitch = (cols + 127) / 128 * 128;
    pool->n_slots = n_slots;
    if (int rc = arena.alloc(&pool->data, pool->slot_bytes() * n_slots))
        return rc;
    if (int rc = encode_pool_map(&pool->tmap_score, *pool, kScoreBoxW, kScoreBoxH))
        return rc;
    return encode_pool_map(&pool->tmap_patch, *pool, kPatchW, kPatchH);
}

TileGrid make_tile_grid(int img_w, int img_h, int cell)
{
    TileGrid g;
    g.cell = cell;
    g.img_w = img_w;
    g.img_h = img_h;
    g.ny = 1 + ((img_h - 1) / cell);
    g.nx = 1 + ((img_w - 1) / cell);
    return g;
}

int make_detect_workspace(DetectWorkspace *ws, DeviceArena &arena, const TileGrid &grid, int rows, int pitch, int batch)
{
    ws->batch = batch;
    ws->n_tiles = grid.count();
    const long tw = std::min(grid.cell, grid.img_w), th = std::min(grid.cell, grid.img_h);
    long cap = 1;
    while (cap < (tw * th) / 2 + 2)
        cap <<= 1; // no two survivors are 4-adjacent: at most half the pixels, rounded to a power of two
    ws->tile_cap = (int)cap;
    const size_t nt = (size_t)batch * ws->n_tiles;
    int rc = arena.alloc(&ws->score, (size_t)batch * rows * pitch);
    rc = rc ? rc : arena.alloc(&ws->parent, (size_t)batch * rows * pitch);
    rc = rc ? rc : arena.alloc(&ws->tile_list, nt * cap);
    rc = rc ? rc : arena.alloc(&ws->tile_aux, nt * cap * 2);
    rc = rc ? rc : arena.alloc(&ws->tile_out, nt * cap);
    rc = rc ? rc : arena.alloc(&ws->tile_count, nt);
    rc = rc ? rc : arena.alloc(&ws->tile_overflow, nt);
    rc = rc ? rc : arena.alloc(&ws->big_list, nt * 64);
    rc = rc ? rc : arena.alloc(&ws->tile_out_count, nt);
    rc = rc ? rc : arena.alloc(&ws->tiles_done, (size_t)batch);
    rc = rc ? rc : arena.alloc(&ws->cand_count, (size_t)batch);
    rc = rc ? rc : arena.alloc(&ws->retry, (size_t)batch);
    rc = rc ? rc : arena.alloc(&ws->error, 1);
    return rc;
}

int make_feat(FeatDev *f, DeviceArena &arena, int cap, int n_cells, int rows)
{
    f->cap = cap;
    int rc = arena.alloc(&f->n, 1);
    rc = rc ? rc : arena.alloc(&f->xy, (size_t)cap);
    rc = rc ? rc : arena.alloc(&f->resp, (size_t)cap);
    rc = rc ? rc : arena.alloc(&f->desc, (size_t)cap * 8);
    rc = rc ? rc : arena.alloc(&f->matched, (size_t)cap);
    rc = rc ? rc : arena.alloc(&f->depth, (size_t)cap);
    rc = rc ? rc : arena.alloc(&f->cell_start, (size_t)n_cells + 1);
    rc = rc ? rc : arena.alloc(&f->cell_items, (size_t)cap);
    rc = rc ? rc : arena.alloc(&f->row_start, (size_t)rows + 2);
    rc = rc ? rc : arena.alloc(&f->row_items, (size_t)cap);
    return rc;
}

int make_points(PointStore *p, DeviceArena &arena, int cap)
{
    p->cap = cap;
    int rc = arena.alloc(&p->xyz, (size_t)cap * 3);
    rc = rc ? rc : arena.alloc(&p->desc, (size_t)cap * 8);
    rc = rc ? rc : arena.alloc(&p->counter, (size_t)cap);
    rc = rc ? rc : arena.alloc(&p->age, (size_t)cap);
    rc = rc ? rc : arena.alloc(&p->match_idx, (size_t)cap);
    return rc;
}

// cv::undistortPoints for one point (host; image bounds only, lvt/src/lvt_local_map.cpp:95-122)
static void undistort_host(const lvt_params_c &p, float u, float v, float *ou, float *ov)
{
    const double fx = p.fx, fy = p.fy, cx = p.cx, cy = p.cy;
    double x = ((double)u - cx) / fx, y = ((double)v - cy) / fy;
    const double x0 = x, y0 = y;
    for (int j = 0; j < 5; j++)
    {
        const double r2 = x * x + y * y;
        const double icdist = 1.0 / (1 + (((double)p.k3 * r2 + (double)p.k2) * r2 + (double)p.k1) * r2);
        const double dX = 2 * (double)p.p1 * x * y + (double)p.p2 * (r2 + 2 * x * x);
        const double dY = (double)p.p1 * (r2 + 2 * y * y) + 2 * (double)p.p2 * x * y;
        x = (x0 - dX) * icdist;
        y = (y0 - dY) * icdist;
    }
    *ou = (float)(x * fx + cx);
    *ov = (float)(y * fy + cy);
}

CamParams make_cam_params(const lvt_params_c &p)
{
    CamParams c;
    c.fx = p.fx;
    c.fy = p.fy;
    c.cx = p.cx;
    c.cy = p.cy;
    c.baseline = p.baseline;
    c.near_plane = p.near_plane_distance;
    c.far_plane = p.far_plane_distance;
    if (std::fabs((double)p.k1) < 1e-5)
    {
        c.min_x = 0.0f;
        c.max_x = (float)p.img_width;
        c.min_y = 0.0f;
        c.max_y = (float)p.img_height;
    }
    else
    {
        float x[4], y[4];
        undistort_host(p, 0.0f, 0.0f, &x[0], &y[0]);
        undistort_host(p, (float)p.img_width, 0.0f, &x[1], &y[1]);
        undistort_host(p, 0.0f, (float)p.img_height, &x[2], &y[2]);
        undistort_host(p, (float)p.img_width, (float)p.img_height, &x[3], &y[3]);
        c.min_x = std::min(x[0], x[2]);
        c.max_x = std::max(x[1], x[3]);
        c.min_y = std::min(y[0], y[1]);
        c.max_y = std::max(y[2], y[3]);
    }
    c.img_w = p.img_width;
    c.img_h = p.img_height;
    const float k_cell = (float)kHashCell;
    c.cells_x = (int)std::ceil(p.img_width / k_cell);
    c.cells_y = (int)std::ceil(p.img_height / k_cell);
    c.tracking_radius = p.tracking_radius;
    c.cell_search_radius = (p.tracking_radius == kHashCell) ? 1 : (int)std::ceil((float)p.tracking_radius / k_cell);
    c.tracking_ratio_th = p.tracking_ratio_test_threshold;
    c.triangulation_ratio_th = p.triangulation_ratio_test_threshold;
    c.desc_dist_th = p.descriptor_matching_threshold;
    return c;
}

// BRIEF test pairs (lvt_set_brief_pairs): the process-wide choice and its version.  Every context
// owns a device copy of the table and refreshes it (stream-ordered, on its own device, from its own
// thread) when the version has moved on -- nothing device-side is shared between contexts.
static std::mutex g_pairs_mu;
static signed char g_pairs[256][4];
static bool g_pairs_custom = false;
static std::atomic<unsigned> g_pairs_version{1};

} // namespace lvtb

using namespace lvtb;

// ---------------------------------------------------------------------------------------------
// the context
// ---------------------------------------------------------------------------------------------
struct lvtk_ctx
{
    lvt_params_c params;
    CamParams cam;
    TrackParams tp;
    DetectParams dp;
    int device = 0;
    cudaStream_t stream = nullptr;
    DeviceArena arena;
    ImagePool pool;
    DetectWorkspace ws;
    int fcap = 0, pcap = 0; // feature capacity per image (fixed) / capacity of the point stores (grows, ctx_grow_points)
    int in_cap = 0;         // capacity of the seam / external-corner input buffers (= the initial pcap)
    int n_grown = 0;        // times the point stores were doubled
    static constexpr int kXStreams = 4; // extraction pipelines in flight (lvt_track_pool)
    static constexpr int kSets = 6;     // feature sets (left, right) cycling through them
    FeatDev feats_h[2 * kSets];
    int last_set = 0; // set holding the features of the last tracked frame
    const FeatDev *last_feats_h = nullptr; // host copies of the (left, right) feature sets of the last tracked frame
    FeatDev *feats_d = nullptr;
    int *d_slots = nullptr;
    uint8_t *d_packed = nullptr; // one tightly packed image per pool slot: where a page-locked caller buffer lands (one DMA)
    cudaEvent_t ev_caller_read = nullptr; // the DMA engine is through with the caller's (page-locked) right image
    uint8_t *h_stage = nullptr; // pinned, one image per pool slot, the pool's pitch
    UploadLanes lanes;          // host staging lanes of the blocking entry points
    int upload_bands = 1;       // row bands per image
    int upload_dmas = 1;        // DMAs per image
    float *d_depth = nullptr, *h_depth = nullptr;
    // seam inputs
    float2 *d_in_xy = nullptr;
    float *d_in_resp = nullptr;
    int *d_in_n = nullptr;
    int *d_int_a = nullptr, *d_int_b = nullptr; // [pcap] generic int outputs
    float *d_f_a = nullptr, *d_f_b = nullptr;   // [pcap]
    PoseD *d_pose_out = nullptr;
    // tracking
    TrackState *d_state = nullptr;
    uint8_t *d_ctl = nullptr; // 2 x FrameCtl: hand-over between the kernels of the tracking chain; frames alternate, so
                              // that a frame's early parts read the previous frame's block while that frame's map
                              // maintenance is still using it (ctl_idx: the block of the frame launched last)
    int ctl_idx = 0;
    unsigned frame_seq = 0; // frames launched with an overlap so far (TrackState::rest_seq; wraps)
    cudaEvent_t ev_pose_done[2] = {}, ev_rest_done[2] = {}; // batched engine: TrackOverlap events, by frame parity
    FrameResult *d_result = nullptr, *h_result = nullptr;
    int *h_error = nullptr; // pinned copy of ws.error, fetched together with the result
    EarlyResult *h_early = nullptr, *d_early = nullptr; // pose + state in mapped pinned memory (d_early: its device
                                                        // address), written by the pose solver itself
    int early_seq = 0;
    cudaEvent_t ev_pose = nullptr;      // h_early is valid
    cudaEvent_t ev_frame = nullptr;     // the whole frame (map maintenance, h_result, h_error) is through
    int parity = 0;                     // blocking stereo frames alternate between two sets of buffers
    // lvt_set_rectification: the images of lvt_track / lvt_pool_upload are raw; they land in `raw`
    // (same geometry as the pool) and rectify_kernel writes the pool slot
    bool rectify = false;
    RectifyDev rect[2];
    uint8_t *raw = nullptr;
    cudaEvent_t ev_tl[4] = {};          // LVT_B200_TIMELINE: call start, left image in HBM, left features, pose
    double tl_ms[3] = {0, 0, 0};
    long tl_n = 0;
    // resident frame pool + pipelined streaming (lvt_pool_* / lvt_track_pool)
    cudaStream_t xs[kXStreams] = {}; // extraction runs here, tracking on `stream`
    DetectWorkspace wsx[kXStreams]; // wsx[0] == ws
    cudaEvent_t ev_extracted[kSets] = {}, ev_tracked[kSets] = {};
    cudaEvent_t ev_left = nullptr, ev_right = nullptr; // blocking stereo path: left features indexed / right side ready
    cudaEvent_t ev_batch[2] = {nullptr, nullptr}; // device time of the last lvt_track_pool call
    float last_batch_ms = 0.f;
    DeviceArena rarena;
    ImagePool rpool;
    int rpool_frames = 0;
    int *d_slot_table = nullptr;       // [2 * rpool_frames] = 0, 1, 2, ...

    float *rdepth = nullptr;           // resident RGB-D pool: [rpool_frames][rows][cols] metres

    // ---- the batched engine (lvt_track_pool / lvt_track_batch*): frames go through feature extraction in
    // groups of `G` (one launch of every extraction kernel per group: grid z indexes the images) on one of
    // three extraction streams while the tracking stream consumes their feature sets in order.  Allocated
    // at the first batched call (ctx_ensure_engine).
    struct Engine
    {
        static constexpr int kSlots = 3; // groups in flight: being uploaded / extracted / tracked
        bool ready = false;
        int G = 4;
        DeviceArena arena;
        std::vector<FeatDev> feats_h; // [2 * kSlots * G]
        FeatDev *feats_d = nullptr;
        std::vector<CandLists> row_cand; // [kSlots * G]
        DetectWorkspace ws[kSlots];      // batch = 2 G images
        cudaEvent_t ev_extracted[kSlots] = {}, ev_tracked[kSlots] = {};
        // ring of device image slots for frames that arrive in host memory, their pinned staging, slot table
        ImagePool pool;
        uint8_t *raw = nullptr, *h_stage = nullptr, *d_packed = nullptr;
        int *d_slots = nullptr;
        float *d_depth = nullptr, *h_depth = nullptr;
        // per-frame results of a batch
        FrameResult *d_results = nullptr, *h_results = nullptr;
        int results_cap = 0;
    } eng;

    PointStore map, staged;
    TrackScratch sc;
    CandLists row_cand[kSets]; // per feature set
    TrackLaunchCfg tcfg;       // shared-memory layout of the tracking kernels for this context's capacities
    uint32_t *d_brief_offsets = nullptr, *h_brief_offsets = nullptr; // [8 * 32] device copy / pinned source
    unsigned pairs_version = 0;
};

// one frame's tracking kernels (parts: 1 up to the pose, 2 the rest, 3 both) on `st`
static int ctx_launch_track(lvtk_ctx *c, FrameResult *result, const FeatDev *feats, const CandLists &row_cand, cudaStream_t st,
                            cudaEvent_t right_ready = nullptr, int parts = 3, EarlyResult *early = nullptr, int early_seq = 0,
                            int overlap = 0 /* 0 none, 1 the rest on the side stream, 2 + early map pass */,
                            const FeatDev *next_feats = nullptr /* overlap: the next frame's left features (device), if extracted */)
{
    if (parts & 1)
        c->ctl_idx ^= 1; // a new frame (parts == 2 finishes the frame launched with parts == 1)
    const size_t cb = (frame_ctl_bytes() + 255) & ~(size_t)255;
    uint8_t *ctl = c->d_ctl + cb * (size_t)c->ctl_idx, *ctl_prev = c->d_ctl + cb * (size_t)(c->ctl_idx ^ 1);
    if (overlap)
        c->frame_seq++;
    TrackOverlap ov{c->xs[3], c->ev_pose_done[c->ctl_idx], c->ev_rest_done[c->ctl_idx], c->ev_rest_done[c->ctl_idx ^ 1], overlap == 2,
                    (int)c->frame_seq, next_feats};
    return launch_track_frame(c->d_state, ctl, result, c->map, c->staged, feats, c->tp, c->sc, row_cand, c->tcfg, c->ws.error, st,
                              right_ready, parts, early, early_seq, overlap ? &ov : nullptr, ctl_prev);
}

// The reference's map and staged-point vectors grow without bound (lvt/src/lvt_local_map.cpp:331-353).
// Here a frame that could overflow the stores is refused on the device before it touches anything
// (TrackState::halt, track_a_kernel); the host then doubles every array sized by the point capacity,
// keeps the contents, clears the halt and runs the frame again.  Rare (the default capacity holds
// 8x the feature capacity), so it simply idles the device first.
static int ctx_grow_points(lvtk_ctx *c)
{
    LVT_CUDA_TRY(cudaDeviceSynchronize());
    const size_t o = (size_t)c->pcap, n = 2 * o;
    if (n > ((size_t)1 << 26))
    {
        set_last_error(__FILE__, __LINE__, "point stores cannot grow beyond 2^26 points");
        return LVTK_ERR_CAPACITY;
    }
    DeviceArena &A = c->arena;
    int rc = LVTK_OK;
    for (PointStore *ps : {&c->map, &c->staged})
    {
        rc = rc ? rc : A.regrow(&ps->xyz, o * 3, n * 3);
        rc = rc ? rc : A.regrow(&ps->desc, o * 8, n * 8);
        rc = rc ? rc : A.regrow(&ps->counter, o, n);
        rc = rc ? rc : A.regrow(&ps->age, o, n);
        rc = rc ? rc : A.regrow(&ps->match_idx, o, n);
        ps->cap = (int)n;
    }
    rc = rc ? rc : A.regrow(&c->sc.ms.proj, 0, n); // scratch: nothing to keep
    rc = rc ? rc : A.regrow(&c->sc.ms.vis, 0, n);
    rc = rc ? rc : A.regrow(&c->sc.ms.choice, 0, n);
    rc = rc ? rc : A.regrow(&c->sc.ms.items, 0, 2 * n + 1024);
    rc = rc ? rc : A.regrow(&c->sc.sol_xyz, 0, n * 3);
    rc = rc ? rc : A.regrow(&c->sc.sol_uv, 0, n);
    rc = rc ? rc : A.regrow(&c->sc.level, 0, n);
    rc = rc ? rc : A.regrow(&c->sc.inlier, 0, n);
    rc = rc ? rc : A.regrow(&c->sc.e2, 0, n);
    rc = rc ? rc : A.regrow(&c->sc.map_cand.keys, 0, n * kMapCandCap);
    rc = rc ? rc : A.regrow(&c->sc.map_cand.count, 0, n);
    rc = rc ? rc : A.regrow(&c->sc.bs.proj, 0, n);
    rc = rc ? rc : A.regrow(&c->sc.bs.vis, 0, n);
    rc = rc ? rc : A.regrow(&c->sc.bs.choice, 0, n);
    rc = rc ? rc : A.regrow(&c->sc.bs.items, 0, 2 * n + 1024);
    rc = rc ? rc : A.regrow(&c->sc.staged_cand.keys, 0, n * kMapCandCap);
    rc = rc ? rc : A.regrow(&c->sc.staged_cand.count, 0, n);
    if (rc)
        return rc;
    c->pcap = (int)n;
    LVT_CUDA_TRY(cudaDeviceSynchronize()); // the arena's memsets / copies ran on the default stream
    if (int r2 = launch_clear_halt(c->d_state, c->stream))
        return r2;
    LVT_CUDA_TRY(cudaStreamSynchronize(c->stream));
    c->n_grown++;
    return LVTK_OK;
}

static int ctx_ensure_engine(lvtk_ctx *c)
{
    lvtk_ctx::Engine &E = c->eng;
    if (E.ready)
        return LVTK_OK;
    LVT_CUDA_TRY(cudaDeviceSynchronize());
    E.G = 4;
    if (const char *e = std::getenv("LVT_B200_GROUP")) // frames per extraction launch (tuning aid)
        E.G = std::max(1, std::min(8, std::atoi(e)));
    const int n_sets = lvtk_ctx::Engine::kSlots * E.G, rows = c->params.img_height, cols = c->params.img_width;
    const int n_cells = c->cam.cells_x * c->cam.cells_y;
    E.feats_h.resize(2 * (size_t)n_sets);
    E.row_cand.resize((size_t)n_sets);
    int rc = LVTK_OK;
    for (size_t i = 0; i < E.feats_h.size() && !rc; i++)
        rc = make_feat(&E.feats_h[i], E.arena, c->fcap, n_cells, rows);
    rc = rc ? rc : E.arena.alloc(&E.feats_d, E.feats_h.size());
    for (int i = 0; i < n_sets && !rc; i++)
    {
        rc = E.arena.alloc(&E.row_cand[i].keys, (size_t)c->fcap * kRowCandCap);
        rc = rc ? rc : E.arena.alloc(&E.row_cand[i].count, (size_t)c->fcap);
        E.row_cand[i].cap = kRowCandCap;
    }
    for (int k = 0; k < lvtk_ctx::Engine::kSlots && !rc; k++)
    {
        rc = make_detect_workspace(&E.ws[k], E.arena, c->dp.grid, rows, c->pool.pitch, 2 * E.G);
        E.ws[k].error = c->ws.error; // one sticky error flag per context
    }
    // the ring for host frames: 2 images per frame (RGB-D uses every other slot's worth: G per group)
    rc = rc ? rc : make_image_pool(&E.pool, E.arena, rows, cols, 2 * n_sets);
    rc = rc ? rc : E.arena.alloc(&E.raw, E.pool.slot_bytes() * 2 * (size_t)n_sets);
    rc = rc ? rc : E.arena.alloc(&E.d_packed, (size_t)rows * cols * 2 * (size_t)n_sets);
    rc = rc ? rc : E.arena.alloc(&E.d_slots, 2 * (size_t)n_sets);
    rc = rc ? rc : E.arena.alloc(&E.d_depth, (size_t)n_sets * rows * cols);
    if (rc)
        return rc;
    LVT_CUDA_TRY(cudaMemcpy(E.feats_d, E.feats_h.data(), sizeof(FeatDev) * E.feats_h.size(), cudaMemcpyHostToDevice));
    std::vector<int> table(2 * (size_t)n_sets);
    for (size_t i = 0; i < table.size(); i++)
        table[i] = (int)i;
    LVT_CUDA_TRY(cudaMemcpy(E.d_slots, table.data(), sizeof(int) * table.size(), cudaMemcpyHostToDevice));
    LVT_CUDA_TRY(cudaMallocHost(&E.h_stage, E.pool.slot_bytes() * 2 * (size_t)n_sets));
    std::memset(E.h_stage, 0, E.pool.slot_bytes() * 2 * (size_t)n_sets);
    LVT_CUDA_TRY(cudaMallocHost(&E.h_depth, sizeof(float) * (size_t)n_sets * rows * cols));
    for (int k = 0; k < lvtk_ctx::Engine::kSlots; k++)
    {
        LVT_CUDA_TRY(cudaEventCreateWithFlags(&E.ev_extracted[k], cudaEventDisableTiming));
        LVT_CUDA_TRY(cudaEventCreateWithFlags(&E.ev_tracked[k], cudaEventDisableTiming));
    }
    LVT_CUDA_TRY(cudaDeviceSynchronize()); // the arena's memsets ran on the default stream
    E.ready = true;
    return LVTK_OK;
}

static int ctx_ensure_results(lvtk_ctx *c, int n)
{
    lvtk_ctx::Engine &E = c->eng;
    if (n <= E.results_cap)
        return LVTK_OK;
    LVT_CUDA_TRY(cudaDeviceSynchronize());
    if (E.d_results)
        cudaFree(E.d_results);
    if (E.h_results)
        cudaFreeHost(E.h_results);
    E.d_results = nullptr, E.h_results = nullptr, E.results_cap = 0;
    int cap = 64;
    while (cap < n)
        cap <<= 1;
    LVT_CUDA_TRY(cudaMalloc(&E.d_results, sizeof(FrameResult) * (size_t)cap));
    LVT_CUDA_TRY(cudaMallocHost(&E.h_results, sizeof(FrameResult) * (size_t)cap));
    E.results_cap = cap;
    return LVTK_OK;
}

static void ctx_free_engine(lvtk_ctx *c)
{
    lvtk_ctx::Engine &E = c->eng;
    E.arena.release();
    for (int k = 0; k < lvtk_ctx::Engine::kSlots; k++)
    {
        if (E.ev_extracted[k])
            cudaEventDestroy(E.ev_extracted[k]);
        if (E.ev_tracked[k])
            cudaEventDestroy(E.ev_tracked[k]);
    }
    if (E.h_stage)
        cudaFreeHost(E.h_stage);
    if (E.h_depth)
        cudaFreeHost(E.h_depth);
    if (E.d_results)
        cudaFree(E.d_results);
    if (E.h_results)
        cudaFreeHost(E.h_results);
    E.ready = false;
}

// true when `p` is page-locked host memory (cudaMallocHost / cudaHostRegister / lvt_alloc_pinned): the DMA
// engine reads it directly, the staging copy is skipped
static bool is_pinned_host(const void *p)
{
    cudaPointerAttributes at;
    if (cudaPointerGetAttributes(&at, p) != cudaSuccess)
    {
        cudaGetLastError();
        return false;
    }
    return at.type == cudaMemoryTypeHost;
}

// the context's copy of the BRIEF table follows lvt_set_brief_pairs; `stream` = where the next
// brief_kernel of this context will run
static int ctx_sync_pairs(lvtk_ctx *c, cudaStream_t stream)
{
    if (c->pairs_version == g_pairs_version.load(std::memory_order_acquire))
        return LVTK_OK;
    // the pinned source may still be read by an earlier refresh on another stream
    LVT_CUDA_TRY(cudaDeviceSynchronize());
    {
        std::lock_guard<std::mutex> lk(g_pairs_mu);
        make_brief_offsets(g_pairs_custom ? g_pairs : nullptr, c->h_brief_offsets);
        c->pairs_version = g_pairs_version.load(std::memory_order_relaxed);
    }
    LVT_CUDA_TRY(cudaMemcpyAsync(c->d_brief_offsets, c->h_brief_offsets, sizeof(uint32_t) * 256, cudaMemcpyHostToDevice, stream));
    LVT_CUDA_TRY(cudaStreamSynchronize(stream)); // every stream of the context sees the new table
    return LVTK_OK;
}

static int ctx_check_error(lvtk_ctx *c)
{
    int e = 0;
    LVT_CUDA_TRY(cudaMemcpyAsync(&e, c->ws.error, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
    LVT_CUDA_TRY(cudaStreamSynchronize(c->stream));
    if (e)
    {
        cudaMemsetAsync(c->ws.error, 0, sizeof(int), c->stream);
        set_last_error(__FILE__, __LINE__, "device-side capacity error");
    }
    return e;
}

static int ctx_build(lvtk_ctx *c, const lvt_params_c &p, int device, int n_slots)
{
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0)
    {
        set_last_error(__FILE__, __LINE__, "no CUDA device");
        return LVTK_ERR_NO_DEVICE;
    }
    if (p.img_width <= 0 || p.img_height <= 0 || p.img_width > 4095 || p.img_height > 4095 || p.detection_cell_size <= 0 ||
        p.max_keypoints_per_cell <= 0 || p.tracking_radius <= 0 || p.agast_threshold <= 0 || p.agast_threshold > 254)
    {
        set_last_error(__FILE__, __LINE__, "parameters out of range");
        return LVTK_ERR_ARG;
    }
    if (device >= 0)
        LVT_CUDA_TRY(cudaSetDevice(device));
    LVT_CUDA_TRY(cudaGetDevice(&c->device));
    c->params = p;
    c->cam = make_cam_params(p);
    if (c->cam.cells_x * c->cam.cells_y + p.img_height + 3 > 12000)
    {
        set_last_error(__FILE__, __LINE__, "image too large for the shared-memory hash grid");
        return LVTK_ERR_ARG;
    }
    c->tp.cam = c->cam;
    c->tp.sensor = 1;
    c->tp.min_matches = p.min_num_matches_for_tracking;
    c->tp.untracked_threshold = p.untracked_threshold;
    c->tp.staged_threshold = p.staged_threshold;
    c->tp.triangulation_policy = p.triangulation_policy;
    // the tracking chain is the serial part of a frame: its kernels go in front of the (state-free)
    // extraction of later frames / of the right image whenever an SM frees up
    int prio_lo = 0, prio_hi = 0;
    LVT_CUDA_TRY(cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi));
    if (std::getenv("LVT_B200_NO_PRIORITY"))
        prio_hi = prio_lo;
    LVT_CUDA_TRY(cudaStreamCreateWithPriority(&c->stream, cudaStreamNonBlocking, prio_hi));
    // xs[0..2]: extraction of the batched engine / of the right image; xs[3]: the left (gray) image of a blocking
    // call, which is on the way to the pose
    for (int i = 0; i < lvtk_ctx::kXStreams; i++)
        LVT_CUDA_TRY(cudaStreamCreateWithPriority(&c->xs[i], cudaStreamNonBlocking, i == 3 ? prio_hi : prio_lo));
    for (int i = 0; i < lvtk_ctx::kSets; i++)
    {
        LVT_CUDA_TRY(cudaEventCreateWithFlags(&c->ev_extracted[i], cudaEventDisableTiming));
        LVT_CUDA_TRY(cudaEventCreateWithFlags(&c->ev_tracked[i], cudaEventDisableTiming));
    }
    for (int i = 0; i < 2; i++)
        LVT_CUDA_TRY(cudaEventCreate(&c->ev_batch[i]));
    LVT_CUDA_TRY(cudaEventCreateWithFlags(&c->ev_left, cudaEventDisableTiming));
    LVT_CUDA_TRY(cudaEventCreateWithFlags(&c->ev_right, cudaEventDisableTiming));
    LVT_CUDA_TRY(cudaEventCreateWithFlags(&c->ev_pose, cudaEventDisableTiming));
    for (int k = 0; k < 2; k++)
    {
        LVT_CUDA_TRY(cudaEventCreateWithFlags(&c->ev_pose_done[k], cudaEventDisableTiming));
        LVT_CUDA_TRY(cudaEventCreateWithFlags(&c->ev_rest_done[k], cudaEventDisableTiming));
    }
    LVT_CUDA_TRY(cudaEventCreateWithFlags(&c->ev_frame, cudaEventDisableTiming));

    if (int rc = make_image_pool(&c->pool, c->arena, p.img_height, p.img_width, n_slots))
        return rc;
    if (int rc = c->arena.alloc(&c->raw, c->pool.slot_bytes() * n_slots))
        return rc;
    if (int rc = c->arena.alloc(&c->d_packed, (size_t)p.img_height * p.img_width * n_slots))
        return rc;
    LVT_CUDA_TRY(cudaEventCreateWithFlags(&c->ev_caller_read, cudaEventDisableTiming));
    c->dp.grid = make_tile_grid(p.img_width, p.img_height, p.detection_cell_size);
    c->dp.threshold = p.agast_threshold;
    c->dp.threshold_low = (int)((double)p.agast_threshold * 0.5 + 0.5);
    c->dp.max_per_cell = p.max_keypoints_per_cell;
    c->dp.pitch = c->pool.pitch;
    c->dp.rows = p.img_height;
    c->dp.cols = p.img_width;
    if (c->dp.grid.count() > 1024)
    {
        set_last_error(__FILE__, __LINE__, "more than 1024 detection tiles");
        return LVTK_ERR_ARG;
    }
    if (int rc = make_detect_workspace(&c->ws, c->arena, c->dp.grid, p.img_height, c->pool.pitch, 2))
        return rc;
    c->wsx[0] = c->ws;
    for (int i = 1; i < lvtk_ctx::kXStreams; i++)
    {
        if (int rc = make_detect_workspace(&c->wsx[i], c->arena, c->dp.grid, p.img_height, c->pool.pitch, 2))
            return rc;
        c->wsx[i].error = c->ws.error; // one sticky error flag per context
    }

    // capacities: ANMS keeps >= k+1 per tile (ties add a few); 2x headroom, at least 4096
    long want = 2L * c->dp.grid.count() * (p.max_keypoints_per_cell + 1);
    long fcap = 4096;
    while (fcap < want)
        fcap <<= 1;
    if (fcap > 24576)
        fcap = 24576; // two owner arrays must fit in 227 KB of shared memory
    c->fcap = (int)fcap;
    c->pcap = std::max(32768, 8 * c->fcap);
    if (const char *e = std::getenv("LVT_B200_POINT_CAP")) // test aid: start small, exercise the growth path
        c->pcap = std::max(256, std::atoi(e));
    c->in_cap = std::max(c->pcap, 2 * c->fcap);
    const int n_cells = c->cam.cells_x * c->cam.cells_y;
    for (int i = 0; i < 2 * lvtk_ctx::kSets; i++)
        if (int rc = make_feat(&c->feats_h[i], c->arena, c->fcap, n_cells, p.img_height))
            return rc;
    int rc = c->arena.alloc(&c->feats_d, 2 * lvtk_ctx::kSets);
    rc = rc ? rc : c->arena.alloc(&c->d_slots, 4);
    rc = rc ? rc : c->arena.alloc(&c->d_depth, (size_t)2 * p.img_width * p.img_height); // two alternating frames
    rc = rc ? rc : c->arena.alloc(&c->d_in_xy, (size_t)2 * c->in_cap);
    rc = rc ? rc : c->arena.alloc(&c->d_in_resp, (size_t)2 * c->in_cap);
    rc = rc ? rc : c->arena.alloc(&c->d_in_n, 2);
    rc = rc ? rc : c->arena.alloc(&c->d_int_a, (size_t)c->in_cap);
    rc = rc ? rc : c->arena.alloc(&c->d_int_b, (size_t)c->in_cap);
    rc = rc ? rc : c->arena.alloc(&c->d_f_a, (size_t)c->in_cap);
    rc = rc ? rc : c->arena.alloc(&c->d_f_b, (size_t)c->in_cap);
    rc = rc ? rc : c->arena.alloc(&c->d_pose_out, 1);
    rc = rc ? rc : c->arena.alloc(&c->d_state, 1);
    rc = rc ? rc : c->arena.alloc(&c->d_ctl, 2 * ((frame_ctl_bytes() + 255) & ~(size_t)255));
    rc = rc ? rc : c->arena.alloc(&c->d_result, 1);
    rc = rc ? rc : make_points(&c->map, c->arena, c->pcap);
    rc = rc ? rc : make_points(&c->staged, c->arena, c->pcap);
    rc = rc ? rc : c->arena.alloc(&c->sc.ms.proj, (size_t)c->pcap);
    rc = rc ? rc : c->arena.alloc(&c->sc.ms.vis, (size_t)c->pcap);
    rc = rc ? rc : c->arena.alloc(&c->sc.ms.choice, (size_t)c->pcap);
    rc = rc ? rc : c->arena.alloc(&c->sc.ms.items, (size_t)2 * c->pcap + 1024); // + the per-CTA rounding of the team's work lists
    rc = rc ? rc : c->arena.alloc(&c->sc.sol_xyz, (size_t)c->pcap * 3);
    rc = rc ? rc : c->arena.alloc(&c->sc.sol_uv, (size_t)c->pcap);
    rc = rc ? rc : c->arena.alloc(&c->sc.level, (size_t)c->pcap);
    rc = rc ? rc : c->arena.alloc(&c->sc.inlier, (size_t)c->pcap);
    rc = rc ? rc : c->arena.alloc(&c->sc.e2, (size_t)c->pcap);
    rc = rc ? rc : c->arena.alloc(&c->sc.map_cand.keys, (size_t)c->pcap * kMapCandCap);
    rc = rc ? rc : c->arena.alloc(&c->sc.map_cand.count, (size_t)c->pcap);
    c->sc.map_cand.cap = kMapCandCap;
    rc = rc ? rc : c->arena.alloc(&c->sc.bs.proj, (size_t)c->pcap);
    rc = rc ? rc : c->arena.alloc(&c->sc.bs.vis, (size_t)c->pcap);
    rc = rc ? rc : c->arena.alloc(&c->sc.bs.choice, (size_t)c->pcap);
    rc = rc ? rc : c->arena.alloc(&c->sc.bs.items, (size_t)2 * c->pcap + 1024);
    rc = rc ? rc : c->arena.alloc(&c->sc.staged_cand.keys, (size_t)c->pcap * kMapCandCap);
    rc = rc ? rc : c->arena.alloc(&c->sc.staged_cand.count, (size_t)c->pcap);
    c->sc.staged_cand.cap = kMapCandCap;
    for (int i = 0; i < lvtk_ctx::kSets; i++)
    {
        rc = rc ? rc : c->arena.alloc(&c->row_cand[i].keys, (size_t)c->fcap * kRowCandCap);
        rc = rc ? rc : c->arena.alloc(&c->row_cand[i].count, (size_t)c->fcap);
        c->row_cand[i].cap = kRowCandCap;
    }
    rc = rc ? rc : c->arena.alloc(&c->sc.row_choice, (size_t)c->in_cap);
    rc = rc ? rc : c->arena.alloc(&c->sc.pair_query, (size_t)c->in_cap);
    rc = rc ? rc : c->arena.alloc(&c->sc.pair_train, (size_t)c->in_cap);
    rc = rc ? rc : c->arena.alloc(&c->sc.tri_xyz, (size_t)c->in_cap * 3);
    rc = rc ? rc : c->arena.alloc(&c->sc.tri_ok, (size_t)c->in_cap);
    if (rc)
        return rc;
    LVT_CUDA_TRY(cudaMemcpy(c->feats_d, c->feats_h, sizeof(c->feats_h), cudaMemcpyHostToDevice));
    const int slots[4] = {0, 1, 2, 3};
    LVT_CUDA_TRY(cudaMemcpy(c->d_slots, slots, sizeof(slots), cudaMemcpyHostToDevice));
    LVT_CUDA_TRY(cudaMallocHost(&c->h_stage, (size_t)n_slots * c->pool.pitch * p.img_height));
    std::memset(c->h_stage, 0, (size_t)n_slots * c->pool.pitch * p.img_height);
    LVT_CUDA_TRY(cudaMallocHost(&c->h_depth, sizeof(float) * (size_t)2 * p.img_width * p.img_height));
    LVT_CUDA_TRY(cudaMallocHost(&c->h_result, sizeof(FrameResult)));
    LVT_CUDA_TRY(cudaMallocHost(&c->h_error, sizeof(int)));
    LVT_CUDA_TRY(cudaHostAlloc(&c->h_early, sizeof(EarlyResult), cudaHostAllocMapped));
    std::memset(c->h_early, 0, sizeof(EarlyResult));
    LVT_CUDA_TRY(cudaHostGetDevicePointer(reinterpret_cast<void **>(&c->d_early), c->h_early, 0));
    {
        // staging lanes (the calling thread + helpers, upload.cuh): LVT_B200_UPLOAD_THREADS; row bands
        // per image: LVT_B200_UPLOAD_BANDS.  Default: up to 4 lanes, a quarter of the host cores
        // shared between the GPUs of the box, one band per lane (tools/probe/host_probe.py).
        int ndev_all = 1;
        cudaGetDeviceCount(&ndev_all);
        const int cores = (int)std::thread::hardware_concurrency();
        int lanes = std::max(1, std::min(4, cores / (4 * std::max(1, ndev_all))));
        if (const char *e = std::getenv("LVT_B200_UPLOAD_THREADS"))
            lanes = std::max(1, std::min(16, std::atoi(e)));
        c->upload_bands = lanes > 1 ? 2 * lanes : 1;
        if (const char *e = std::getenv("LVT_B200_UPLOAD_BANDS"))
            c->upload_bands = std::max(1, std::min(16, std::atoi(e)));
        c->upload_dmas = 1; // one DMA per image measured best (tools/probe/timeline_probe.py)
        if (const char *e = std::getenv("LVT_B200_UPLOAD_DMAS"))
            c->upload_dmas = std::max(1, std::min(16, std::atoi(e)));
        c->lanes.start(c->device, lanes - 1);
    }
    if (int r2 = c->arena.alloc(&c->d_brief_offsets, 256))
        return r2;
    LVT_CUDA_TRY(cudaMallocHost(&c->h_brief_offsets, sizeof(uint32_t) * 256));
    if (int r2 = track_configure(c->fcap, &c->tcfg))
        return r2;
    LVT_CUDA_TRY(cudaDeviceSynchronize()); // the arena's memsets ran on the default stream
    if (int r2 = ctx_sync_pairs(c, c->stream))
        return r2;
    if (int r3 = launch_reset_state(c->d_state, c->stream))
        return r3;
    LVT_CUDA_TRY(cudaStreamSynchronize(c->stream));
    return LVTK_OK;
}

static void ctx_free(lvtk_ctx *c)
{
    if (!c)
        return;
    cudaSetDevice(c->device);
    c->lanes.stop();
    if (c->stream)
    {
        cudaStreamSynchronize(c->stream);
        cudaStreamDestroy(c->stream);
    }
    for (int i = 0; i < lvtk_ctx::kXStreams; i++)
        if (c->xs[i])
        {
            cudaStreamSynchronize(c->xs[i]);
            cudaStreamDestroy(c->xs[i]);
        }
    for (int i = 0; i < lvtk_ctx::kSets; i++)
    {
        if (c->ev_extracted[i])
            cudaEventDestroy(c->ev_extracted[i]);
        if (c->ev_tracked[i])
            cudaEventDestroy(c->ev_tracked[i]);
    }
    for (int i = 0; i < 2; i++)
        if (c->ev_batch[i])
            cudaEventDestroy(c->ev_batch[i]);
    if (c->ev_left)
        cudaEventDestroy(c->ev_left);
    if (c->ev_right)
        cudaEventDestroy(c->ev_right);
    if (c->ev_pose)
        cudaEventDestroy(c->ev_pose);
    for (int k = 0; k < 2; k++)
    {
        if (c->ev_pose_done[k])
            cudaEventDestroy(c->ev_pose_done[k]);
        if (c->ev_rest_done[k])
            cudaEventDestroy(c->ev_rest_done[k]);
    }
    if (c->ev_frame)
        cudaEventDestroy(c->ev_frame);
    if (c->ev_caller_read)
        cudaEventDestroy(c->ev_caller_read);
    if (c->h_early)
        cudaFreeHost(c->h_early);
    c->rarena.release();
    ctx_free_engine(c);
    c->arena.release();
    if (c->h_stage)
        cudaFreeHost(c->h_stage);
    if (c->h_depth)
        cudaFreeHost(c->h_depth);
    if (c->h_result)
        cudaFreeHost(c->h_result);
    if (c->h_error)
        cudaFreeHost(c->h_error);
    if (c->h_brief_offsets)
        cudaFreeHost(c->h_brief_offsets);
    delete c;
}

// host image (any stride) -> pinned staging -> pitched pool slot, in row bands (upload.cuh);
// ctx_stage_image queues an image, ctx_stage_flush stages and enqueues everything queued
static void ctx_stage_image(lvtk_ctx *c, int slot, const uint8_t *img, int rows, int cols, int stride, bool to_raw = false)
{
    c->lanes.add_image(img, (size_t)stride, c->h_stage + (size_t)slot * rows * c->pool.pitch,
                       (to_raw ? c->raw : c->pool.data) + c->pool.slot_bytes() * slot, c->pool.pitch, (size_t)cols, rows,
                       c->upload_bands, c->upload_dmas);
}

// raw image in c->raw[slot] -> rectified image in the pool slot, as camera `cam` sees it
static int ctx_rectify_slot(lvtk_ctx *c, int slot, int cam, cudaStream_t stream)
{
    const uint8_t *raw[2] = {c->raw + c->pool.slot_bytes() * slot, nullptr};
    uint8_t *dst[2] = {c->pool.data + c->pool.slot_bytes() * slot, nullptr};
    const RectifyDev *cams[2] = {&c->rect[cam], nullptr};
    return launch_rectify(raw, dst, 1, cams, c->pool, stream);
}

// a page-locked caller image -> pool slot: one contiguous DMA + re-pitch on the device, no staging copy
static int ctx_upload_pinned(lvtk_ctx *c, int slot, const uint8_t *img, cudaStream_t stream, bool to_raw = false)
{
    const int rows = c->params.img_height, cols = c->params.img_width;
    uint8_t *packed = c->d_packed + (size_t)rows * cols * (size_t)slot;
    LVT_CUDA_TRY(cudaMemcpyAsync(packed, img, (size_t)rows * cols, cudaMemcpyHostToDevice, stream));
    return launch_repitch(packed, (to_raw ? c->raw : c->pool.data) + c->pool.slot_bytes() * slot, rows, cols, c->pool.pitch, stream);
}

static int ctx_stage_flush(lvtk_ctx *c, cudaStream_t stream = nullptr)
{
    LVT_CUDA_TRY(c->lanes.run(stream ? stream : c->stream));
    return LVTK_OK;
}

static int ctx_upload_image(lvtk_ctx *c, int slot, const uint8_t *img, int rows, int cols, int stride)
{
    ctx_stage_image(c, slot, img, rows, cols, stride);
    return ctx_stage_flush(c);
}

static int ctx_download_features(lvtk_ctx *c, int which, lvtk_keypoint *out_kps, uint8_t *out_desc, int cap, int *n_out)
{
    const FeatDev &f = c->feats_h[which];
    int n = 0;
    LVT_CUDA_TRY(cudaMemcpyAsync(&n, f.n, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
    LVT_CUDA_TRY(cudaStreamSynchronize(c->stream));
    *n_out = n;
    if (n > cap)
        return LVTK_ERR_CAPACITY;
    if (n == 0)
        return LVTK_OK;
    std::vector<float2> xy(n);
    std::vector<float> resp(n);
    LVT_CUDA_TRY(cudaMemcpyAsync(xy.data(), f.xy, sizeof(float2) * n, cudaMemcpyDeviceToHost, c->stream));
    LVT_CUDA_TRY(cudaMemcpyAsync(resp.data(), f.resp, sizeof(float) * n, cudaMemcpyDeviceToHost, c->stream));
    if (out_desc)
        LVT_CUDA_TRY(cudaMemcpyAsync(out_desc, f.desc, (size_t)32 * n, cudaMemcpyDeviceToHost, c->stream));
    LVT_CUDA_TRY(cudaStreamSynchronize(c->stream));
    if (out_kps)
        for (int i = 0; i < n; i++)
            out_kps[i] = lvtk_keypoint{xy[i].x, xy[i].y, resp[i]};
    return LVTK_OK;
}

// keypoints + descriptors + marks of a seam call -> feature set `which`, indexed
static int ctx_upload_features(lvtk_ctx *c, int which, const lvtk_keypoint *kps, const uint8_t *desc, int n,
                               const uint8_t *marks)
{
    if (n > c->fcap)
        return LVTK_ERR_CAPACITY;
    const FeatDev &f = c->feats_h[which];
    std::vector<float2> xy(n);
    std::vector<float> resp(n);
    for (int i = 0; i < n; i++)
    {
        xy[i] = make_float2(kps[i].x, kps[i].y);
        resp[i] = kps[i].response;
    }
    LVT_CUDA_TRY(cudaMemcpyAsync(f.n, &n, sizeof(int), cudaMemcpyHostToDevice, c->stream));
    if (n)
    {
        LVT_CUDA_TRY(cudaMemcpyAsync(f.xy, xy.data(), sizeof(float2) * n, cudaMemcpyHostToDevice, c->stream));
        LVT_CUDA_TRY(cudaMemcpyAsync(f.resp, resp.data(), sizeof(float) * n, cudaMemcpyHostToDevice, c->stream));
        LVT_CUDA_TRY(cudaMemcpyAsync(f.desc, desc, (size_t)32 * n, cudaMemcpyHostToDevice, c->stream));
    }
    LVT_CUDA_TRY(cudaStreamSynchronize(c->stream)); // the host vectors go out of scope
    if (int rc = launch_index(c->feats_d + which, 1, c->cam, c->stream))
        return rc;
    if (marks && n)
        LVT_CUDA_TRY(cudaMemcpyAsync(f.matched, marks, n, cudaMemcpyHostToDevice, c->stream));
    return LVTK_OK;
}

static PoseD make_pose(const double q[4], const double t[3])
{
    PoseD p;
    p.q = Quat{q[0], q[1], q[2], q[3]};
    p.t[0] = t[0];
    p.t[1] = t[1];
    p.t[2] = t[2];
    return p;
}

// ---------------------------------------------------------------------------------------------
// lvt_system mirror
// ---------------------------------------------------------------------------------------------
namespace
{
struct System
{
    lvtk_ctx *ctx = nullptr;
    int sensor = 1;
    int state = 1;
    int frame_number = 0;
    PoseD last_pose;
    lvt_frame_info info;
    bool pending = false; // the last blocking frame returned at its pose; info / error flag are still on their way
    int last_status = LVTK_OK; // of the last tracking call on this handle (lvt_get_last_status)

    // host-side time of the blocking entry points: [0] staging + H2D enqueue, [1] kernel enqueue,
    // [2] waiting for the device, [3] calls  (lvt_debug_host_times)
    double host_us[4] = {0, 0, 0, 0};
    std::chrono::steady_clock::time_point marks[4];
    void host_mark(int k)
    {
        marks[k] = std::chrono::steady_clock::now();
        if (k == 3)
        {
            for (int i = 0; i < 3; i++)
                host_us[i] += std::chrono::duration<double, std::micro>(marks[i + 1] - marks[i]).count();
            host_us[3] += 1;
        }
    }

    System()
    {
        last_pose.q = Quat{1, 0, 0, 0};
        last_pose.t[0] = last_pose.t[1] = last_pose.t[2] = 0;
        std::memset(&info, 0, sizeof(info));
        info.state = 1;
    }

    // lvt_track hands the pose back as soon as the solver is through; the rest of that frame (staged
    // points, triangulation, counters) is collected here, before anything needs it
    int finish_pending()
    {
        if (!pending)
            return LVTK_OK;
        pending = false;
        lvtk_ctx *c = ctx;
        LVT_CUDA_TRY(cudaEventSynchronize(c->ev_frame));
        if (const int e = *c->h_error)
        {
            cudaMemsetAsync(c->ws.error, 0, sizeof(int), c->stream);
            set_last_error(__FILE__, __LINE__, "device-side capacity error");
            return e;
        }
        info = c->h_result->info;
        info.frame_number = frame_number;
        return LVTK_OK;
    }

    // the LOST short-circuit of lvt_system::track (lvt/src/lvt_system.cpp:159-166)
    bool lost_shortcut(PoseD *out)
    {
        if (state == 3)
            finish_pending();
        frame_number++;
        if (state != 3)
            return false;
        const int map_n = info.map_points_after, staged_n = info.staged_after; // the map is left as it was
        std::memset(&info, 0, sizeof(info));
        info.frame_number = frame_number;
        info.state = 3;
        info.map_points_after = map_n;
        info.staged_after = staged_n;
        *out = last_pose;
        return true;
    }

    // features of this frame are in ctx->feats[0..1]: run perform_tracking on the device
    int run_tracking(PoseD *out)
    {
        lvtk_ctx *c = ctx;
        c->last_set = 0;
        c->last_feats_h = c->feats_h;
        if (int rc = launch_index(c->feats_d, sensor == 1 ? 2 : 1, c->cam, c->stream))
            return rc;
        if (sensor == 1)
            if (int rc = launch_rowcand(c->feats_d, c->cam, c->row_cand[0], c->stream))
                return rc;
        for (;;)
        {
            if (int rc = ctx_launch_track(c, c->d_result, c->feats_d, c->row_cand[0], c->stream))
                return rc;
            if (int rc = fetch_result(out))
                return rc;
            if (c->h_result->info.state != 0)
                return LVTK_OK;
            if (int rc = ctx_grow_points(c)) // the frame was refused: the point stores could have overflowed
                return rc;
        }
    }

    // the frame's kernels are enqueued on ctx->stream: fetch the result (the one host<->device round trip)
    int fetch_result(PoseD *out)
    {
        lvtk_ctx *c = ctx;
        LVT_CUDA_TRY(cudaMemcpyAsync(c->h_result, c->d_result, sizeof(FrameResult), cudaMemcpyDeviceToHost, c->stream));
        LVT_CUDA_TRY(cudaMemcpyAsync(c->h_error, c->ws.error, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
        host_mark(2);
        LVT_CUDA_TRY(cudaStreamSynchronize(c->stream));
        host_mark(3);
        pending = false;
        if (const int e = *c->h_error)
        {
            cudaMemsetAsync(c->ws.error, 0, sizeof(int), c->stream);
            set_last_error(__FILE__, __LINE__, "device-side capacity error");
            return e;
        }
        if (c->h_result->info.state == 0)
            return LVTK_OK; // refused (TrackState::halt): the caller grows the point stores and runs the frame again
        info = c->h_result->info;
        info.frame_number = frame_number;
        state = info.state;
        *out = c->h_result->pose;
        return LVTK_OK;
    }

    // A blocking frame came back refused (h_early->state == 0): its features are extracted and indexed,
    // only the tracking kernels run again after the point stores have grown.
    int rerun_refused(const FeatDev *feats, int s)
    {
        lvtk_ctx *c = ctx;
        while (c->h_early->state == 0)
        {
            if (int rc = ctx_grow_points(c)) // idles the device first
                return rc;
            cudaStream_t st = c->stream;
            if (int rc = ctx_launch_track(c, c->d_result, feats, c->row_cand[s], st, nullptr, 3, c->d_early, ++c->early_seq))
                return rc;
            LVT_CUDA_TRY(cudaEventRecord(c->ev_pose, st));
            LVT_CUDA_TRY(cudaMemcpyAsync(c->h_result, c->d_result, sizeof(FrameResult), cudaMemcpyDeviceToHost, st));
            LVT_CUDA_TRY(cudaMemcpyAsync(c->h_error, c->ws.error, sizeof(int), cudaMemcpyDeviceToHost, st));
            LVT_CUDA_TRY(cudaEventRecord(c->ev_frame, st));
            if (int rc = wait_for_pose())
                return rc;
        }
        return LVTK_OK;
    }

    // the pose solver stores its result straight into pinned host memory, sequence number last
    int wait_for_pose()
    {
        lvtk_ctx *c = ctx;
        const int want = c->early_seq;
        unsigned spins = 0;
        while (__atomic_load_n(&c->h_early->seq, __ATOMIC_ACQUIRE) != want)
        {
            if ((++spins & 0xFFFu) == 0 && cudaEventQuery(c->ev_pose) != cudaErrorNotReady)
                break;
            LVT_CPU_RELAX();
        }
        if (__atomic_load_n(&c->h_early->seq, __ATOMIC_ACQUIRE) != want)
        {
            LVT_CUDA_TRY(cudaEventSynchronize(c->ev_pose));
            if (__atomic_load_n(&c->h_early->seq, __ATOMIC_ACQUIRE) != want)
            {
                set_last_error(__FILE__, __LINE__, "the pose solver did not report");
                return LVTK_ERR_CUDA;
            }
        }
        return LVTK_OK;
    }

    // Blocking stereo frame.  Only the left image is on the path to the pose (map matching and the
    // solver never look at the right one), so the two images go down two streams: the left one is
    // staged, uploaded and extracted first and the tracking kernels up to the pose solver are queued
    // right behind it; then the host stages the right image, a second stream extracts it and lists
    // the row-matching candidates, and both meet in front of track_b (staged points, triangulation).
    // The call returns when the pose is on the host: track_b and the right image finish while the
    // caller prepares the next frame, everything after them is ordered behind them by the streams.
    int track_stereo(const uint8_t *left, const uint8_t *right, int rows, int cols, PoseD *out)
    {
        lvtk_ctx *c = ctx;
        if (sensor != 1 || rows != c->params.img_height || cols != c->params.img_width)
            return LVTK_ERR_ARG;
        if (int rc = ctx_sync_pairs(c, c->xs[0]))
            return rc;
        if (pending && cudaEventQuery(c->ev_frame) == cudaSuccess)
            if (int rc = finish_pending()) // surfaces an error of the previous frame's map maintenance
                return rc;             // (before the frame counter moves: the new frame is not consumed)
        if (lost_shortcut(out))
            return LVTK_OK;
        const bool first_frame = state == 1;
        // Buffers (pool slots, staging, feature sets, candidate lists) alternate between two sets, so
        // nothing of this frame waits for the previous frame's map maintenance except the map
        // matching itself.  The set used two frames ago is free: its track_b ran before the pose of
        // the previous frame, which this thread has waited for.
        const int s = c->parity;
        c->parity ^= 1;
        cudaStream_t xl = c->xs[3], xr = c->xs[1], st = c->stream;
        FeatDev *feats = c->feats_d + 2 * s;
        const int *slots = c->d_slots + 2 * s;
        host_mark(0);
        c->last_set = s;
        c->last_feats_h = c->feats_h + 2 * s;
        static const bool timeline = std::getenv("LVT_B200_TIMELINE") != nullptr;
        if (timeline)
        {
            if (!c->ev_tl[0])
                for (int k = 0; k < 4; k++)
                    cudaEventCreate(&c->ev_tl[k]);
            cudaEventRecord(c->ev_tl[0], xl);
        }
        const bool pinned_in = is_pinned_host(left) && is_pinned_host(right);
        if (pinned_in)
        {
            if (int rc = ctx_upload_pinned(c, 2 * s, left, xl, c->rectify))
                return rc;
        }
        else
        {
            ctx_stage_image(c, 2 * s, left, rows, cols, cols, c->rectify);
            if (int rc = ctx_stage_flush(c, xl))
                return rc;
        }
        if (timeline)
            cudaEventRecord(c->ev_tl[1], xl);
        if (c->rectify)
            if (int rc = ctx_rectify_slot(c, 2 * s, 0, xl))
                return rc;
        if (int rc = launch_detect(c->pool, c->wsx[0], c->dp, slots, 1, feats, kBriefBorder, 1, xl))
            return rc;
        const bool fused_index = brief_can_index(c->cam);
        if (int rc = launch_brief(c->pool, slots, 1, feats, c->d_brief_offsets, xl, fused_index ? &c->cam : nullptr))
            return rc;
        if (!fused_index)
            if (int rc = launch_index(feats, 1, c->cam, xl))
                return rc;
        LVT_CUDA_TRY(cudaEventRecord(c->ev_left, xl));
        if (timeline)
            cudaEventRecord(c->ev_tl[2], xl);
        LVT_CUDA_TRY(cudaStreamWaitEvent(st, c->ev_left, 0));
        if (int rc = ctx_launch_track(c, c->d_result, feats, c->row_cand[s], st, nullptr, 1, c->d_early, ++c->early_seq))
            return rc;
        LVT_CUDA_TRY(cudaEventRecord(c->ev_pose, st));
        if (timeline)
            cudaEventRecord(c->ev_tl[3], st);
        if (pinned_in)
        {
            if (int rc = ctx_upload_pinned(c, 2 * s + 1, right, xr, c->rectify))
                return rc;
            LVT_CUDA_TRY(cudaEventRecord(c->ev_caller_read, xr)); // the buffers are the caller's again when the call returns
        }
        else
        {
            ctx_stage_image(c, 2 * s + 1, right, rows, cols, cols, c->rectify);
            if (int rc = ctx_stage_flush(c, xr))
                return rc;
        }
        if (c->rectify)
            if (int rc = ctx_rectify_slot(c, 2 * s + 1, 1, xr))
                return rc;
        host_mark(1);
        if (int rc = launch_detect(c->pool, c->wsx[1], c->dp, slots + 1, 1, feats + 1, kBriefBorder, 1, xr))
            return rc;
        if (int rc = launch_brief(c->pool, slots + 1, 1, feats + 1, c->d_brief_offsets, xr, fused_index ? &c->cam : nullptr))
            return rc;
        if (!fused_index)
            if (int rc = launch_index(feats + 1, 1, c->cam, xr))
                return rc;
        LVT_CUDA_TRY(cudaStreamWaitEvent(xr, c->ev_left, 0)); // the candidates pair left with right descriptors
        if (int rc = launch_rowcand(feats, c->cam, c->row_cand[s], xr))
            return rc;
        LVT_CUDA_TRY(cudaEventRecord(c->ev_right, xr));
        if (int rc = ctx_launch_track(c, c->d_result, feats, c->row_cand[s], st, c->ev_right, 2))
            return rc;
        LVT_CUDA_TRY(cudaMemcpyAsync(c->h_result, c->d_result, sizeof(FrameResult), cudaMemcpyDeviceToHost, st));
        LVT_CUDA_TRY(cudaMemcpyAsync(c->h_error, c->ws.error, sizeof(int), cudaMemcpyDeviceToHost, st));
        LVT_CUDA_TRY(cudaEventRecord(c->ev_frame, st));
        host_mark(2);
        if (int rc = wait_for_pose())
            return rc;
        if (c->h_early->state == 0)
            if (int rc = rerun_refused(feats, s))
                return rc;
        if (pinned_in)
            LVT_CUDA_TRY(cudaEventSynchronize(c->ev_caller_read)); // long done: the pose came after the left image's whole chain
        host_mark(3);
        if (timeline)
        {
            cudaEventSynchronize(c->ev_tl[3]);
            for (int k = 0; k < 3; k++)
            {
                float ms = 0;
                cudaEventElapsedTime(&ms, c->ev_tl[k], c->ev_tl[k + 1]);
                c->tl_ms[k] += ms;
            }
            if (++c->tl_n % 50 == 0)
            {
                std::fprintf(stderr, "timeline (mean of 50 frames): stage+H2D %.1f us | left extraction %.1f us | mapcand..pose+copy %.1f us\n",
                             20.0 * c->tl_ms[0], 20.0 * c->tl_ms[1], 20.0 * c->tl_ms[2]);
                c->tl_ms[0] = c->tl_ms[1] = c->tl_ms[2] = 0;
            }
        }
        pending = true;
        const PoseD pose = c->h_early->pose;
        state = c->h_early->state;
        info.state = state;
        info.frame_number = frame_number;
        if (!first_frame && state == 2)
            last_pose = pose; // m_last_pose = computed_pose (lvt_system.cpp:205); not on the first frame
        *out = pose;
        return LVTK_OK;
    }

    int track_external(const uint8_t *left, const uint8_t *right, int rows, int cols, const double (*cl)[2], int nl,
                       const double (*cr)[2], int nr, PoseD *out)
    {
        lvtk_ctx *c = ctx;
        if (sensor != 1 || rows != c->params.img_height || cols != c->params.img_width || nl < 0 || nr < 0)
            return LVTK_ERR_ARG;
        if (nl > c->in_cap || nr > c->in_cap)
            return LVTK_ERR_CAPACITY;
        if (int rc = finish_pending())
            return rc;
        if (int rc = ctx_sync_pairs(c, c->stream))
            return rc;
        if (lost_shortcut(out))
            return LVTK_OK;
        host_mark(0);
        ctx_stage_image(c, 0, left, rows, cols, cols);
        ctx_stage_image(c, 1, right, rows, cols, cols);
        if (int rc = ctx_stage_flush(c))
            return rc;
        host_mark(1);
        std::vector<float2> xy((size_t)2 * c->in_cap);
        for (int i = 0; i < nl; i++)
            xy[i] = make_float2((float)cl[i][0], (float)cl[i][1]);
        for (int i = 0; i < nr; i++)
            xy[c->in_cap + i] = make_float2((float)cr[i][0], (float)cr[i][1]);
        const int n_in[2] = {nl, nr};
        LVT_CUDA_TRY(cudaMemcpyAsync(c->d_in_xy, xy.data(), sizeof(float2) * xy.size(), cudaMemcpyHostToDevice, c->stream));
        LVT_CUDA_TRY(cudaMemcpyAsync(c->d_in_n, n_in, sizeof(n_in), cudaMemcpyHostToDevice, c->stream));
        LVT_CUDA_TRY(cudaStreamSynchronize(c->stream));
        if (int rc = launch_border_filter(c->d_in_xy, nullptr, c->d_in_n, c->in_cap, c->feats_d, 2, rows, cols, c->ws.error,
                                          c->stream))
            return rc;
        if (int rc = launch_brief(c->pool, c->d_slots, 2, c->feats_d, c->d_brief_offsets, c->stream))
            return rc;
        return finish(out);
    }

    // Blocking RGB-D frame, the same shape as track_stereo: the gray image is staged, uploaded and
    // extracted first; the depth image (4 bytes per pixel, the bulk of the upload) is staged and uploaded
    // on a second stream meanwhile and is needed only by the depth gate behind the descriptors; the call
    // returns at the pose, the map maintenance finishes behind it on alternating buffer sets.
    int track_rgbd(const uint8_t *gray, const float *depth, int rows, int cols, PoseD *out)
    {
        lvtk_ctx *c = ctx;
        if (sensor != 2 || rows != c->params.img_height || cols != c->params.img_width)
            return LVTK_ERR_ARG;
        if (int rc = ctx_sync_pairs(c, c->xs[0]))
            return rc;
        if (pending && cudaEventQuery(c->ev_frame) == cudaSuccess)
            if (int rc = finish_pending())
                return rc;
        if (lost_shortcut(out))
            return LVTK_OK;
        const bool first_frame = state == 1;
        const int s = c->parity;
        c->parity ^= 1;
        cudaStream_t xl = c->xs[3], xr = c->xs[1], st = c->stream;
        FeatDev *feats = c->feats_d + 2 * s;
        const int *slots = c->d_slots + 2 * s;
        const size_t npx = (size_t)rows * cols;
        float *h_depth = c->h_depth + s * npx, *d_depth = c->d_depth + s * npx;
        host_mark(0);
        c->last_set = s;
        c->last_feats_h = c->feats_h + 2 * s;
        const bool pinned_in = is_pinned_host(gray) && is_pinned_host(depth);
        if (pinned_in)
        {
            if (int rc = ctx_upload_pinned(c, 2 * s, gray, xl))
                return rc;
        }
        else
        {
            ctx_stage_image(c, 2 * s, gray, rows, cols, cols);
            if (int rc = ctx_stage_flush(c, xl))
                return rc;
        }
        if (int rc = launch_detect(c->pool, c->wsx[0], c->dp, slots, 1, feats, kBriefBorder, 1, xl))
            return rc;
        if (int rc = launch_brief(c->pool, slots, 1, feats, c->d_brief_offsets, xl))
            return rc;
        if (pinned_in)
        {
            LVT_CUDA_TRY(cudaMemcpyAsync(d_depth, depth, sizeof(float) * npx, cudaMemcpyHostToDevice, xr));
            LVT_CUDA_TRY(cudaEventRecord(c->ev_caller_read, xr));
        }
        else
        {
            c->lanes.add_image(depth, sizeof(float) * (size_t)cols, h_depth, d_depth, sizeof(float) * (size_t)cols,
                               sizeof(float) * (size_t)cols, rows, 2 * c->upload_bands, 2);
            if (int rc = ctx_stage_flush(c, xr))
                return rc;
        }
        LVT_CUDA_TRY(cudaEventRecord(c->ev_right, xr));
        host_mark(1);
        LVT_CUDA_TRY(cudaStreamWaitEvent(xl, c->ev_right, 0)); // the depth image has arrived
        if (int rc = launch_depth_gate(c->feats_h[2 * s], d_depth, c->params, xl))
            return rc;
        if (int rc = launch_index(feats, 1, c->cam, xl))
            return rc;
        LVT_CUDA_TRY(cudaEventRecord(c->ev_left, xl));
        LVT_CUDA_TRY(cudaStreamWaitEvent(st, c->ev_left, 0));
        if (int rc = ctx_launch_track(c, c->d_result, feats, c->row_cand[s], st, nullptr, 1, c->d_early, ++c->early_seq))
            return rc;
        LVT_CUDA_TRY(cudaEventRecord(c->ev_pose, st));
        if (int rc = ctx_launch_track(c, c->d_result, feats, c->row_cand[s], st, nullptr, 2))
            return rc;
        LVT_CUDA_TRY(cudaMemcpyAsync(c->h_result, c->d_result, sizeof(FrameResult), cudaMemcpyDeviceToHost, st));
        LVT_CUDA_TRY(cudaMemcpyAsync(c->h_error, c->ws.error, sizeof(int), cudaMemcpyDeviceToHost, st));
        LVT_CUDA_TRY(cudaEventRecord(c->ev_frame, st));
        host_mark(2);
        if (int rc = wait_for_pose())
            return rc;
        if (c->h_early->state == 0)
            if (int rc = rerun_refused(feats, s))
                return rc;
        if (pinned_in)
            LVT_CUDA_TRY(cudaEventSynchronize(c->ev_caller_read));
        host_mark(3);
        pending = true;
        const PoseD pose = c->h_early->pose;
        state = c->h_early->state;
        info.state = state;
        info.frame_number = frame_number;
        if (!first_frame && state == 2)
            last_pose = pose;
        *out = pose;
        return LVTK_OK;
    }

    // ---- resident pool: frames already in HBM -------------------------------------------------------
    int pool_reserve(int n_frames)
    {
        lvtk_ctx *c = ctx;
        if (n_frames <= 0)
            return LVTK_ERR_ARG;
        finish_pending();
        LVT_CUDA_TRY(cudaDeviceSynchronize());
        c->rarena.release();
        c->rpool_frames = 0;
        c->rdepth = nullptr;
        const int per = sensor == 1 ? 2 : 1;
        if (int rc = make_image_pool(&c->rpool, c->rarena, c->params.img_height, c->params.img_width, per * n_frames))
            return rc;
        int rc = c->rarena.alloc(&c->d_slot_table, (size_t)per * n_frames);
        if (sensor == 2)
            rc = rc ? rc : c->rarena.alloc(&c->rdepth, (size_t)n_frames * c->params.img_height * c->params.img_width);
        if (rc)
            return rc;
        std::vector<int> table((size_t)per * n_frames);
        for (size_t i = 0; i < table.size(); i++)
            table[i] = (int)i;
        LVT_CUDA_TRY(cudaMemcpy(c->d_slot_table, table.data(), sizeof(int) * table.size(), cudaMemcpyHostToDevice));
        LVT_CUDA_TRY(cudaDeviceSynchronize());
        c->rpool_frames = n_frames;
        return LVTK_OK;
    }

    // stereo: (left, right); RGB-D: (gray, nullptr, depth)
    int pool_upload(int frame, const uint8_t *left, const uint8_t *right, const float *depth)
    {
        lvtk_ctx *c = ctx;
        if (frame < 0 || frame >= c->rpool_frames || !left || (sensor == 1 && !right) || (sensor == 2 && !depth))
            return LVTK_ERR_ARG;
        const int per = sensor == 1 ? 2 : 1, rows = c->params.img_height, cols = c->params.img_width;
        const uint8_t *src[2] = {left, right};
        if (int rc = finish_pending())
            return rc;
        for (int k = 0; k < per; k++)
        {
            uint8_t *slot = c->rpool.data + c->rpool.slot_bytes() * (size_t)(per * frame + k);
            uint8_t *dst = c->rectify && sensor == 1 ? c->raw + c->pool.slot_bytes() * k : slot;
            LVT_CUDA_TRY(cudaMemcpy2DAsync(dst, c->rpool.pitch, src[k], cols, cols, rows, cudaMemcpyHostToDevice, c->stream));
        }
        if (sensor == 2)
            LVT_CUDA_TRY(cudaMemcpyAsync(c->rdepth + (size_t)frame * rows * cols, depth, sizeof(float) * (size_t)rows * cols,
                                         cudaMemcpyHostToDevice, c->stream));
        if (c->rectify && sensor == 1)
        {
            // raw frames are rectified once, on their way into the resident pool
            const uint8_t *raw[2] = {c->raw, c->raw + c->pool.slot_bytes()};
            uint8_t *dst[2] = {c->rpool.data + c->rpool.slot_bytes() * (size_t)(2 * frame),
                               c->rpool.data + c->rpool.slot_bytes() * (size_t)(2 * frame + 1)};
            const RectifyDev *cams[2] = {&c->rect[0], &c->rect[1]};
            if (int rc = launch_rectify(raw, dst, 2, cams, c->rpool, c->stream))
                return rc;
        }
        LVT_CUDA_TRY(cudaStreamSynchronize(c->stream));
        return LVTK_OK;
    }

    // ---- the batched engine ---------------------------------------------------------------------------
    // Where the frames of a batch come from: the resident pool (first + i) or the caller's host memory.
    struct FrameSource
    {
        int first = 0;                       // resident: pool frame of batch frame 0
        const uint8_t *const *a = nullptr;   // host: left / gray images, tightly packed
        const uint8_t *const *b = nullptr;   // host: right images (stereo)
        const float *const *depth = nullptr; // host: depth images in metres (RGB-D)
        bool host() const { return a != nullptr; }
    };

    // one host image -> ring slot `slot` on stream sx: page-locked memory is read by the DMA engine directly,
    // pageable memory goes through the pinned staging slot (row bands copied by the staging lanes)
    int upload_ring_image(const uint8_t *img, int slot, cudaStream_t sx, bool to_raw)
    {
        lvtk_ctx *c = ctx;
        lvtk_ctx::Engine &E = c->eng;
        const int rows = c->params.img_height, cols = c->params.img_width;
        uint8_t *dst = (to_raw ? E.raw : E.pool.data) + E.pool.slot_bytes() * (size_t)slot;
        if (is_pinned_host(img))
        {
            // one contiguous DMA of the packed image, re-pitched on the device (a pitched 2-D copy is 3.7x slower)
            uint8_t *packed = E.d_packed + (size_t)rows * cols * (size_t)slot;
            LVT_CUDA_TRY(cudaMemcpyAsync(packed, img, (size_t)rows * cols, cudaMemcpyHostToDevice, sx));
            return launch_repitch(packed, dst, rows, cols, E.pool.pitch, sx);
        }
        c->lanes.add_image(img, (size_t)cols, E.h_stage + E.pool.slot_bytes() * (size_t)slot, dst, E.pool.pitch, (size_t)cols,
                           rows, c->upload_bands, c->upload_dmas);
        LVT_CUDA_TRY(c->lanes.run(sx));
        return LVTK_OK;
    }

    // frames [start, n) of a batch: enqueue everything, wait, fetch the results into eng.h_results[start..n)
    int run_frames(const FrameSource &src, int start, int n)
    {
        lvtk_ctx *c = ctx;
        lvtk_ctx::Engine &E = c->eng;
        constexpr int kSlots = lvtk_ctx::Engine::kSlots;
        const int G = E.G, per = sensor == 1 ? 2 : 1;
        const int rows = c->params.img_height, cols = c->params.img_width;
        const size_t npx = (size_t)rows * cols;
        const bool fused_index = sensor == 1 && brief_can_index(c->cam);
        host_mark(0);
        host_mark(1);
        // timed on the device: first upload / extraction launch .. last result copy
        LVT_CUDA_TRY(cudaEventRecord(c->ev_batch[0], c->stream));
        for (int k = 0; k < kSlots; k++)
            LVT_CUDA_TRY(cudaStreamWaitEvent(c->xs[k], c->ev_batch[0], 0));
        // group gi: frames g0 .. g0 + gn - 1, extraction stream / workspace / feature sets of slot gi % 3.
        // Extraction is state-free (lvt_image_features_handler.cpp:196-209): up to three groups are in
        // flight (uploading, extracting, being tracked); a slot is reused once its frames have been tracked.
        const int n_groups = (n - start + G - 1) / G;
        auto enqueue_extraction = [&](int gi) -> int {
            const int g0 = start + gi * G;
            const int gn = std::min(G, n - g0), slot = gi % kSlots, set0 = slot * G;
            cudaStream_t sx = c->xs[slot];
            if (gi >= kSlots)
            {
                if (src.host()) // the slot's staging buffers and ring images were last read by that extraction
                    LVT_CUDA_TRY(cudaEventSynchronize(E.ev_extracted[slot]));
                LVT_CUDA_TRY(cudaStreamWaitEvent(sx, E.ev_tracked[slot], 0));
            }
            const ImagePool *pool = &c->rpool;
            const int *slots = c->d_slot_table + (size_t)per * (src.first + g0);
            const float *depth = sensor == 2 ? c->rdepth + npx * (size_t)(src.first + g0) : nullptr;
            if (src.host())
            {
                pool = &E.pool;
                slots = E.d_slots + per * set0;
                for (int k = 0; k < gn; k++)
                {
                    const bool raw = c->rectify && sensor == 1;
                    if (int rc = upload_ring_image(src.a[g0 + k], per * (set0 + k), sx, raw))
                        return rc;
                    if (sensor == 1)
                    {
                        if (int rc = upload_ring_image(src.b[g0 + k], per * (set0 + k) + 1, sx, raw))
                            return rc;
                        if (raw)
                        {
                            const size_t sb = E.pool.slot_bytes();
                            const uint8_t *rw[2] = {E.raw + sb * (size_t)(2 * (set0 + k)), E.raw + sb * (size_t)(2 * (set0 + k) + 1)};
                            uint8_t *dst[2] = {E.pool.data + sb * (size_t)(2 * (set0 + k)), E.pool.data + sb * (size_t)(2 * (set0 + k) + 1)};
                            const RectifyDev *cams[2] = {&c->rect[0], &c->rect[1]};
                            if (int rc = launch_rectify(rw, dst, 2, cams, E.pool, sx))
                                return rc;
                        }
                    }
                    else
                    {
                        float *dd = E.d_depth + npx * (size_t)(set0 + k);
                        if (is_pinned_host(src.depth[g0 + k]))
                            LVT_CUDA_TRY(cudaMemcpyAsync(dd, src.depth[g0 + k], sizeof(float) * npx, cudaMemcpyHostToDevice, sx));
                        else
                        {
                            c->lanes.add_image(src.depth[g0 + k], sizeof(float) * (size_t)cols, E.h_depth + npx * (size_t)(set0 + k), dd,
                                               sizeof(float) * (size_t)cols, sizeof(float) * (size_t)cols, rows, 2 * c->upload_bands, 2);
                            LVT_CUDA_TRY(c->lanes.run(sx));
                        }
                    }
                }
                depth = sensor == 2 ? E.d_depth + npx * (size_t)set0 : nullptr;
            }
            // extraction: ONE launch of every kernel for the per * gn images of the group
            FeatDev *feats = E.feats_d + per * set0;
            const FeatDev *feats_h = E.feats_h.data() + per * set0;
            if (int rc = launch_detect(*pool, E.ws[slot], c->dp, slots, per * gn, feats, kBriefBorder, 1, sx))
                return rc;
            if (int rc = launch_brief(*pool, slots, per * gn, feats, c->d_brief_offsets, sx, fused_index ? &c->cam : nullptr))
                return rc;
            if (sensor == 2)
                for (int k = 0; k < gn; k++)
                    if (int rc = launch_depth_gate(feats_h[k], depth + npx * (size_t)k, c->params, sx))
                        return rc;
            if (!fused_index)
                if (int rc = launch_index(feats, per * gn, c->cam, sx))
                    return rc;
            if (sensor == 1)
                for (int k = 0; k < gn; k++)
                    if (int rc = launch_rowcand(feats + 2 * k, c->cam, E.row_cand[set0 + k], sx))
                        return rc;
            LVT_CUDA_TRY(cudaEventRecord(E.ev_extracted[slot], sx));
            return LVTK_OK;
        };
        // LVT_B200_EXTRACT_FIRST (measurement aid, batches of at most three groups): all extraction runs to completion
        // before the first tracking kernel is launched, so the tracking chain is timed with the GPU to itself
        static const bool overlap_on = !std::getenv("LVT_B200_NO_OVERLAP"); // A/B aid: the whole chain on one stream
        static const bool extract_first = std::getenv("LVT_B200_EXTRACT_FIRST") != nullptr;
        const bool alone = extract_first && n_groups <= kSlots;
        if (alone)
        {
            for (int gi = 0; gi < n_groups; gi++)
                if (int rc = enqueue_extraction(gi))
                    return rc;
            LVT_CUDA_TRY(cudaDeviceSynchronize());
            LVT_CUDA_TRY(cudaEventRecord(c->ev_batch[0], c->stream));
        }
        else if (int rc = enqueue_extraction(0))
            return rc;
        for (int gi = 0; gi < n_groups; gi++)
        {
            const int g0 = start + gi * G;
            const int gn = std::min(G, n - g0), slot = gi % kSlots, set0 = slot * G;
            // one group ahead: the last frame of this group lists the points it appends for the next group's first
            // frame, so that group's extraction is enqueued before this group's tracking
            const bool more = gi + 1 < n_groups;
            if (!alone && more)
                if (int rc = enqueue_extraction(gi + 1))
                    return rc;
            FeatDev *feats = E.feats_d + per * set0;
            const FeatDev *feats_h = E.feats_h.data() + per * set0;
            // tracking: strictly in order on the tracking stream; a frame's map maintenance (stagedcand + track_b) goes
            // to the side stream, and from the second frame of the call on the map pass starts early, next to the
            // previous frame's map maintenance (TrackOverlap, track_a_kernel)
            LVT_CUDA_TRY(cudaStreamWaitEvent(c->stream, E.ev_extracted[slot], 0));
            for (int k = 0; k < gn; k++)
            {
                const int overlap = !overlap_on ? 0 : (g0 + k > start ? 2 : 1);
                const FeatDev *next = nullptr;
                if (overlap && k + 1 < gn)
                    next = feats + per * (k + 1);
                else if (overlap && more)
                {
                    const int nslot = (gi + 1) % kSlots;
                    LVT_CUDA_TRY(cudaStreamWaitEvent(c->xs[3], E.ev_extracted[nslot], 0));
                    next = E.feats_d + per * (nslot * G);
                }
                if (int rc = ctx_launch_track(c, E.d_results + (g0 + k), feats + per * k, E.row_cand[set0 + k], c->stream, nullptr, 3,
                                              nullptr, 0, overlap, next))
                    return rc;
            }
            // the group's feature sets are free once the map maintenance of its last frame is through
            LVT_CUDA_TRY(cudaEventRecord(E.ev_tracked[slot], overlap_on ? c->xs[3] : c->stream));
            c->last_feats_h = feats_h + per * (gn - 1);
        }
        if (overlap_on) // the results are complete behind the last frame's track_b
            LVT_CUDA_TRY(cudaStreamWaitEvent(c->stream, c->ev_rest_done[c->ctl_idx], 0));
        LVT_CUDA_TRY(cudaMemcpyAsync(E.h_results + start, E.d_results + start, sizeof(FrameResult) * (size_t)(n - start),
                                     cudaMemcpyDeviceToHost, c->stream));
        LVT_CUDA_TRY(cudaEventRecord(c->ev_batch[1], c->stream));
        host_mark(2); // everything is enqueued ([1]: host time of the enqueue loop); from here the host only waits
        LVT_CUDA_TRY(cudaStreamSynchronize(c->stream));
        for (int k = 0; k < kSlots; k++)
            LVT_CUDA_TRY(cudaStreamSynchronize(c->xs[k]));
        host_mark(3);
        float ms = 0.f;
        LVT_CUDA_TRY(cudaEventElapsedTime(&ms, c->ev_batch[0], c->ev_batch[1]));
        c->last_batch_ms += ms;
        if (int e = ctx_check_error(c))
            return e;
        return LVTK_OK;
    }

    // n frames of `src`, in order: exactly what n blocking calls would return
    int track_frames(const FrameSource &src, int n, double *poses /* n x 12: R row-major, t */, lvt_frame_info *infos)
    {
        lvtk_ctx *c = ctx;
        if (int rc = finish_pending())
            return rc;
        if (int rc = ctx_ensure_engine(c))
            return rc;
        if (int rc = ctx_ensure_results(c, n))
            return rc;
        if (int rc = ctx_sync_pairs(c, c->stream))
            return rc;
        // A frame that could overflow the point stores is refused on the device, and so is every frame
        // behind it (TrackState::halt): grow the stores and run the rest of the batch again.
        c->last_batch_ms = 0.f;
        const FrameResult *res = c->eng.h_results;
        for (int start = 0; start < n;)
        {
            if (int rc = run_frames(src, start, n))
                return rc;
            int refused = n;
            for (int i = start; i < n && refused == n; i++)
                if (res[i].info.state == 0)
                    refused = i;
            if (refused == n)
                break;
            if (int rc = ctx_grow_points(c))
                return rc;
            start = refused;
        }
        for (int i = 0; i < n; i++)
        {
            const FrameResult &r = res[i];
            const bool first_frame = state == 1;
            frame_number++;
            info = r.info;
            info.frame_number = frame_number;
            state = info.state;
            if (!first_frame && state == 2)
                last_pose = r.pose;
            if (poses)
            {
                double m[9];
                quat_to_mat(r.pose.q, m);
                std::memcpy(poses + 12 * (size_t)i, m, sizeof(m));
                std::memcpy(poses + 12 * (size_t)i + 9, r.pose.t, 3 * sizeof(double));
            }
            if (infos)
                infos[i] = info;
        }
        return LVTK_OK;
    }

    int track_pool(int first, int n, double *poses, lvt_frame_info *infos)
    {
        if (first < 0 || n <= 0 || first + n > ctx->rpool_frames)
            return LVTK_ERR_ARG;
        FrameSource src;
        src.first = first;
        return track_frames(src, n, poses, infos);
    }

    int track_batch(int n, const uint8_t *const *a, const uint8_t *const *b, const float *const *depth, int rows, int cols,
                    double *poses, lvt_frame_info *infos)
    {
        lvtk_ctx *c = ctx;
        if (n <= 0 || !a || (sensor == 1 && !b) || (sensor == 2 && !depth) || rows != c->params.img_height ||
            cols != c->params.img_width)
            return LVTK_ERR_ARG;
        for (int i = 0; i < n; i++)
            if (!a[i] || (sensor == 1 && !b[i]) || (sensor == 2 && !depth[i]))
                return LVTK_ERR_ARG;
        FrameSource src;
        src.a = a;
        src.b = b;
        src.depth = depth;
        return track_frames(src, n, poses, infos);
    }

    int finish(PoseD *out)
    {
        const bool first_frame = state == 1;
        PoseD pose;
        if (int rc = run_tracking(&pose))
            return rc;
        if (!first_frame && state == 2)
            last_pose = pose; // m_last_pose = computed_pose (lvt_system.cpp:205); not on the first frame
        *out = pose;
        return LVTK_OK;
    }
};

void write_pose(const PoseD &pose, double R[3][3], double t[3])
{
    double m[9];
    quat_to_mat(pose.q, m);
    for (int i = 0; i < 3; i++)
        for (int j = 0; j < 3; j++)
            R[i][j] = m[3 * i + j];
    t[0] = pose.t[0];
    t[1] = pose.t[1];
    t[2] = pose.t[2];
}
} // namespace

extern "C"
{

// ---- reference ABI (lvt/src/lvt_c.cpp:33-148) -------------------------------------------------
LVT_API lvt_handle lvt_create_from_params(const lvt_params_c *p, int sensor_type)
{
    if (!p || !(sensor_type == 1 || sensor_type == 2))
        return nullptr;
    System *vo = new System();
    vo->sensor = sensor_type;
    vo->ctx = new lvtk_ctx();
    if (ctx_build(vo->ctx, *p, -1, 4) != LVTK_OK)
    {
        ctx_free(vo->ctx);
        delete vo;
        return nullptr;
    }
    vo->ctx->tp.sensor = sensor_type;
    return vo;
}

LVT_API lvt_handle lvt_create(const char *config_file_name, int sensor_type)
{
    lvt_params_c p;
    if (!params_from_file(&p, config_file_name))
        return nullptr;
    return lvt_create_from_params(&p, sensor_type);
}

LVT_API void lvt_destroy(lvt_handle h)
{
    System *vo = static_cast<System *>(h);
    if (!vo)
        return;
    if (vo->ctx)
    {
        cudaSetDevice(vo->ctx->device);
        cudaDeviceSynchronize(); // a frame that returned at its pose may still be finishing
    }
    ctx_free(vo->ctx);
    delete vo;
}

LVT_API void lvt_reset(lvt_handle h)
{
    System *vo = static_cast<System *>(h);
    if (!vo)
        return;
    cudaSetDevice(vo->ctx->device);
    vo->finish_pending();
    cudaStreamSynchronize(vo->ctx->xs[1]);
    launch_reset_state(vo->ctx->d_state, vo->ctx->stream);
    cudaStreamSynchronize(vo->ctx->stream);
    lvtk_ctx *c = vo->ctx;
    const int sensor = vo->sensor;
    *vo = System();
    vo->ctx = c;
    vo->sensor = sensor;
}

LVT_API void lvt_track(lvt_handle h, unsigned char *left, unsigned char *right, int n_rows, int n_cols, double R[3][3],
                       double t[3])
{
    System *vo = static_cast<System *>(h);
    if (!vo || !left || !right)
        return;
    cudaSetDevice(vo->ctx->device);
    PoseD pose;
    if ((vo->last_status = vo->track_stereo(left, right, n_rows, n_cols, &pose)) == LVTK_OK)
        write_pose(pose, R, t);
}

LVT_API void lvt_track_rgbd(lvt_handle h, const unsigned char *gray, const float *depth_m, int n_rows, int n_cols,
                            double R[3][3], double t[3])
{
    System *vo = static_cast<System *>(h);
    if (!vo || !gray || !depth_m)
        return;
    cudaSetDevice(vo->ctx->device);
    PoseD pose;
    if ((vo->last_status = vo->track_rgbd(gray, depth_m, n_rows, n_cols, &pose)) == LVTK_OK)
        write_pose(pose, R, t);
}

LVT_API void lvt_track_with_external_corners(lvt_handle h, unsigned char *left, unsigned char *right, int n_rows,
                                             int n_cols, double corners_left[][2], int n_left,
                                             double corners_right[][2], int n_right, double R[3][3], double t[3])
{
    System *vo = static_cast<System *>(h);
    if (!vo || !left || !right)
        return;
    cudaSetDevice(vo->ctx->device);
    PoseD pose;
    if ((vo->last_status = vo->track_external(left, right, n_rows, n_cols, corners_left, n_left, corners_right, n_right,
                                              &pose)) == LVTK_OK)
        write_pose(pose, R, t);
}

LVT_API int lvt_get_status(lvt_handle h)
{
    System *vo = static_cast<System *>(h);
    return vo ? vo->state : -1;
}

LVT_API int lvt_get_last_status(lvt_handle h)
{
    System *vo = static_cast<System *>(h);
    return vo ? vo->last_status : LVTK_ERR_ARG;
}

// ---- extensions --------------------------------------------------------------------------------
LVT_API void lvt_params_default(lvt_params_c *p) { params_default(p); }
LVT_API int lvt_params_from_file(lvt_params_c *p, const char *f) { return params_from_file(p, f); }

LVT_API int lvt_get_frame_info(lvt_handle h, lvt_frame_info *out)
{
    System *vo = static_cast<System *>(h);
    if (!vo || !out)
        return -1;
    cudaSetDevice(vo->ctx->device);
    vo->finish_pending();
    *out = vo->info;
    return 0;
}

LVT_API int lvt_get_last_pose(lvt_handle h, double q[4], double t[3])
{
    System *vo = static_cast<System *>(h);
    if (!vo)
        return -1;
    q[0] = vo->last_pose.q.w;
    q[1] = vo->last_pose.q.x;
    q[2] = vo->last_pose.q.y;
    q[3] = vo->last_pose.q.z;
    t[0] = vo->last_pose.t[0];
    t[1] = vo->last_pose.t[1];
    t[2] = vo->last_pose.t[2];
    return 0;
}

LVT_API int lvt_debug_get_features(lvt_handle h, int which, float *kps_xy, unsigned char *desc, int cap)
{
    System *vo = static_cast<System *>(h);
    if (!vo || which < 0 || which > 1)
        return -1;
    lvtk_ctx *c = vo->ctx;
    cudaSetDevice(c->device);
    vo->finish_pending();
    const FeatDev &f = (c->last_feats_h ? c->last_feats_h : c->feats_h)[which];
    int n = 0;
    if (cudaMemcpy(&n, f.n, sizeof(int), cudaMemcpyDeviceToHost) != cudaSuccess)
        return -1;
    if (vo->frame_number == 0 || (which == 1 && vo->sensor != 1))
        n = 0;
    const int m = std::min(n, cap);
    if (m > 0 && kps_xy)
        cudaMemcpy(kps_xy, f.xy, sizeof(float2) * m, cudaMemcpyDeviceToHost);
    if (m > 0 && desc)
        cudaMemcpy(desc, f.desc, (size_t)32 * m, cudaMemcpyDeviceToHost);
    return n;
}

LVT_API int lvt_debug_get_points(lvt_handle h, int which, double *xyz, unsigned char *desc, int *counters, int *ages,
                                 int *match_idx, int cap)
{
    System *vo = static_cast<System *>(h);
    if (!vo || which < 0 || which > 1)
        return -1;
    lvtk_ctx *c = vo->ctx;
    cudaSetDevice(c->device);
    vo->finish_pending();
    TrackState st;
    if (cudaMemcpy(&st, c->d_state, sizeof(st), cudaMemcpyDeviceToHost) != cudaSuccess)
        return -1;
    const PointStore &p = which ? c->staged : c->map;
    const int n = which ? st.staged_n : st.map_n;
    const int m = std::min(n, cap);
    if (m > 0)
    {
        if (xyz)
            cudaMemcpy(xyz, p.xyz, sizeof(double) * 3 * m, cudaMemcpyDeviceToHost);
        if (desc)
            cudaMemcpy(desc, p.desc, (size_t)32 * m, cudaMemcpyDeviceToHost);
        if (counters)
            cudaMemcpy(counters, p.counter, sizeof(int) * m, cudaMemcpyDeviceToHost);
        if (ages)
            cudaMemcpy(ages, p.age, sizeof(int) * m, cudaMemcpyDeviceToHost);
        if (match_idx)
            cudaMemcpy(match_idx, p.match_idx, sizeof(int) * m, cudaMemcpyDeviceToHost);
    }
    return n;
}

LVT_API int lvt_debug_point_capacity(lvt_handle h)
{
    System *vo = static_cast<System *>(h);
    return vo ? vo->ctx->pcap : -1;
}

LVT_API int lvt_set_brief_pairs(const signed char pairs[256][4])
{
    if (pairs)
    {
        for (int i = 0; i < 256; i++)
            for (int j = 0; j < 4; j++)
                if (pairs[i][j] < -24 || pairs[i][j] > 24)
                    return -1;
    }
    // live contexts (any device, any thread) pick the table up at their next extraction
    std::lock_guard<std::mutex> lk(g_pairs_mu);
    if (pairs)
        std::memcpy(g_pairs, pairs, sizeof(g_pairs));
    g_pairs_custom = pairs != nullptr;
    g_pairs_version.fetch_add(1, std::memory_order_release);
    return 0;
}

LVT_API const char *lvtk_last_error(void) { return last_error(); }

LVT_API int lvt_set_rectification(lvt_handle h, const lvt_rectify_c *left, const lvt_rectify_c *right)
{
    System *vo = static_cast<System *>(h);
    if (!vo || (left == nullptr) != (right == nullptr))
        return LVTK_ERR_ARG;
    cudaSetDevice(vo->ctx->device);
    vo->finish_pending();
    lvtk_ctx *c = vo->ctx;
    c->rectify = false;
    if (left)
    {
        RectifyDev l, r;
        if (make_rectify_dev(*left, &l) != LVTK_OK || make_rectify_dev(*right, &r) != LVTK_OK)
            return LVTK_ERR_ARG;
        c->rect[0] = l;
        c->rect[1] = r;
        c->rectify = true;
    }
    return LVTK_OK;
}

LVT_API int lvt_pool_reserve(lvt_handle h, int n_frames)
{
    System *vo = static_cast<System *>(h);
    if (!vo)
        return LVTK_ERR_ARG;
    cudaSetDevice(vo->ctx->device);
    return vo->pool_reserve(n_frames);
}
LVT_API int lvt_pool_upload(lvt_handle h, int frame, const unsigned char *left, const unsigned char *right)
{
    System *vo = static_cast<System *>(h);
    if (!vo)
        return LVTK_ERR_ARG;
    cudaSetDevice(vo->ctx->device);
    if (vo->sensor != 1)
        return LVTK_ERR_ARG;
    return vo->pool_upload(frame, left, right, nullptr);
}
LVT_API int lvt_pool_upload_rgbd(lvt_handle h, int frame, const unsigned char *gray, const float *depth_m)
{
    System *vo = static_cast<System *>(h);
    if (!vo || vo->sensor != 2)
        return LVTK_ERR_ARG;
    cudaSetDevice(vo->ctx->device);
    return vo->pool_upload(frame, gray, nullptr, depth_m);
}
LVT_API int lvt_track_batch(lvt_handle h, int n_frames, const unsigned char *const *left, const unsigned char *const *right,
                            int n_rows, int n_cols, double *poses, lvt_frame_info *infos)
{
    System *vo = static_cast<System *>(h);
    if (!vo || vo->sensor != 1)
        return LVTK_ERR_ARG;
    cudaSetDevice(vo->ctx->device);
    const int rc = vo->last_status = vo->track_batch(n_frames, left, right, nullptr, n_rows, n_cols, poses, infos);
    if (prof_enabled())
        prof_collect();
    return rc;
}
LVT_API int lvt_track_batch_rgbd(lvt_handle h, int n_frames, const unsigned char *const *gray, const float *const *depth_m,
                                 int n_rows, int n_cols, double *poses, lvt_frame_info *infos)
{
    System *vo = static_cast<System *>(h);
    if (!vo || vo->sensor != 2)
        return LVTK_ERR_ARG;
    cudaSetDevice(vo->ctx->device);
    const int rc = vo->last_status = vo->track_batch(n_frames, gray, nullptr, depth_m, n_rows, n_cols, poses, infos);
    if (prof_enabled())
        prof_collect();
    return rc;
}
LVT_API void *lvt_alloc_pinned(size_t bytes)
{
    void *p = nullptr;
    if (cudaMallocHost(&p, bytes ? bytes : 1) != cudaSuccess)
    {
        cudaGetLastError();
        return nullptr;
    }
    return p;
}
LVT_API void lvt_free_pinned(void *p)
{
    if (p)
        cudaFreeHost(p);
}
LVT_API int lvt_track_pool(lvt_handle h, int first_frame, int n_frames, double *poses, lvt_frame_info *infos)
{
    System *vo = static_cast<System *>(h);
    if (!vo)
        return LVTK_ERR_ARG;
    cudaSetDevice(vo->ctx->device);
    const int rc = vo->last_status = vo->track_pool(first_frame, n_frames, poses, infos);
    if (prof_enabled())
        prof_collect();
    return rc;
}
/* profiling aid: clock64() phase marks and fixed-point round counts of pool frame i of the last batch
 * (or of the last lvt_track call when i < 0); rounds[4] = evaluations of the pose solver in that frame */
LVT_API int lvt_debug_phase_cycles(lvt_handle h, int i, long long cycles[8], int rounds[8])
{
    System *vo = static_cast<System *>(h);
    if (!vo)
        return -1;
    vo->finish_pending();
    const FrameResult &r = i < 0 ? *vo->ctx->h_result : vo->ctx->eng.h_results[i];
    std::memcpy(cycles, r.cycles, sizeof(r.cycles));
    std::memcpy(rounds, r.rounds, sizeof(r.rounds));
    return 0;
}
/* profiling aid: nanosecond marks inside track_b of pool frame i of the last batch (i < 0: the last blocking call):
 * [0] start, [1] staged rounds, [2] promotion, [3] staged compaction, [4] row matching, [5] triangulation,
 * [6] new points appended, [7] state + prediction written; [8], [9] start / end of the frame's early map pass
 * (track_a_kernel part 1, batched engine); 0 = phase not run */
LVT_API int lvt_debug_frame_marks(lvt_handle h, int i, long long marks[12])
{
    System *vo = static_cast<System *>(h);
    if (!vo)
        return -1;
    vo->finish_pending();
    const FrameResult &r = i < 0 ? *vo->ctx->h_result : vo->ctx->eng.h_results[i];
    std::memcpy(marks, r.dbg, sizeof(r.dbg));
    std::memcpy(marks + 8, r.amark, sizeof(r.amark)); // [8] / [9]: start / end of the early map pass (batched engine)
    return 0;
}
/* profiling aid: host time of the blocking calls since the last reset, microseconds:
 * out[0] staging + H2D enqueue, out[1] kernel enqueue, out[2] waiting for the device, out[3] calls */
LVT_API int lvt_debug_host_times(lvt_handle h, double out[4], int reset)
{
    System *vo = static_cast<System *>(h);
    if (!vo)
        return -1;
    std::memcpy(out, vo->host_us, sizeof(vo->host_us));
    if (reset)
        std::memset(vo->host_us, 0, sizeof(vo->host_us));
    return 0;
}
LVT_API double lvt_last_batch_ms(lvt_handle h)
{
    System *vo = static_cast<System *>(h);
    return vo ? (double)vo->ctx->last_batch_ms : -1.0;
}
LVT_API long lvt_launch_count(void) { return launch_count(); }
LVT_API void lvt_set_profiling(int on)
{
    cudaDeviceSynchronize();
    if (g_prof.used)
        prof_collect();
    g_prof.on = on != 0;
}
LVT_API int lvt_get_kernel_times(double *ms, long *counts, int cap)
{
    cudaDeviceSynchronize();
    if (g_prof.used)
        prof_collect();
    for (int i = 0; i < K_COUNT && i < cap; i++)
    {
        ms[i] = g_prof.ms[i];
        counts[i] = g_prof.count[i];
    }
    return K_COUNT;
}
LVT_API void lvt_reset_kernel_times(void)
{
    cudaDeviceSynchronize();
    g_prof.used = 0;
    for (int i = 0; i < K_COUNT; i++)
    {
        g_prof.ms[i] = 0;
        g_prof.count[i] = 0;
    }
}
LVT_API const char *lvt_kernel_name(int id)
{
    static const char *names[K_COUNT] = {"score_kernel",   "nms_tile_kernel", "tile_kernel",    "gather_kernel",
                                         "brief_kernel",   "index_kernel",    "track_a_kernel", "mapcand_kernel",
                                         "rowcand_kernel", "pose_kernel",     "stagedcand_kernel", "track_b_kernel",
                                         "rectify_kernel", "track_a_kernel[early part]", "mapcand_kernel[early]",
                                         "mapcand_kernel[appended points]"};
    return id >= 0 && id < K_COUNT ? names[id] : "";
}

// ---- seam ABI (include/lvt_kernels.h) ------------------------------------------------------------
LVT_API lvtk_ctx *lvtk_ctx_create(const lvt_params_c *p, int device)
{
    if (!p)
        return nullptr;
    lvtk_ctx *c = new lvtk_ctx();
    if (ctx_build(c, *p, device, 2) != LVTK_OK)
    {
        ctx_free(c);
        return nullptr;
    }
    return c;
}
LVT_API void lvtk_ctx_destroy(lvtk_ctx *c) { ctx_free(c); }
LVT_API int lvtk_is_gpu(void) { return 1; }

LVT_API int lvtk_agast(lvtk_ctx *ctx, const uint8_t *img, int rows, int cols, int stride, int threshold, int nonmax,
                       lvtk_keypoint *out, int cap, int *n_out)
{
    if (!ctx || !img || !n_out || rows <= 0 || cols <= 0 || stride < cols || threshold <= 0 || threshold > 254 ||
        rows > 4095 || cols > 4095)
        return LVTK_ERR_ARG;
    cudaSetDevice(ctx->device);
    // one image == one tile of arbitrary size: a throw-away pool / workspace of that shape
    DeviceArena arena;
    ImagePool pool;
    DetectWorkspace ws;
    FeatDev f, *d_f = nullptr;
    int *d_slot = nullptr;
    DetectParams dp;
    dp.grid = make_tile_grid(cols, rows, std::max(rows, cols));
    dp.threshold = dp.threshold_low = threshold;
    dp.max_per_cell = 0x7FFFFFFF;
    int rc = make_image_pool(&pool, arena, rows, cols, 1);
    dp.pitch = pool.pitch;
    dp.rows = rows;
    dp.cols = cols;
    rc = rc ? rc : make_detect_workspace(&ws, arena, dp.grid, rows, pool.pitch, 1);
    rc = rc ? rc : make_feat(&f, arena, std::max(cap, 1), 1, rows);
    rc = rc ? rc : arena.alloc(&d_f, 1);
    rc = rc ? rc : arena.alloc(&d_slot, 1);
    auto body = [&]() -> int {
        if (rc)
            return rc;
        LVT_CUDA_TRY(cudaMemcpy(d_f, &f, sizeof(f), cudaMemcpyHostToDevice));
        LVT_CUDA_TRY(cudaDeviceSynchronize());
        // stream-ordered: a synchronous copy from pageable memory may return before its DMA lands,
        // and the kernels below run on a non-blocking stream
        LVT_CUDA_TRY(cudaMemcpy2DAsync(pool.data, pool.pitch, img, stride, cols, rows, cudaMemcpyHostToDevice, ctx->stream));
        if (int r = launch_detect(pool, ws, dp, d_slot, 1, d_f, 0, nonmax ? 1 : 0, ctx->stream))
            return r;
        LVT_CUDA_TRY(cudaStreamSynchronize(ctx->stream));
        int n = 0, err = 0;
        LVT_CUDA_TRY(cudaMemcpy(&n, f.n, sizeof(int), cudaMemcpyDeviceToHost));
        LVT_CUDA_TRY(cudaMemcpy(&err, ws.error, sizeof(int), cudaMemcpyDeviceToHost));
        *n_out = n;
        if (err)
            return err;
        std::vector<float2> xy(std::max(n, 1));
        std::vector<float> resp(std::max(n, 1));
        if (n)
        {
            LVT_CUDA_TRY(cudaMemcpy(xy.data(), f.xy, sizeof(float2) * n, cudaMemcpyDeviceToHost));
            LVT_CUDA_TRY(cudaMemcpy(resp.data(), f.resp, sizeof(float) * n, cudaMemcpyDeviceToHost));
        }
        for (int i = 0; i < n; i++)
            out[i] = lvtk_keypoint{xy[i].x, xy[i].y, resp[i]};
        return LVTK_OK;
    };
    rc = body();
    arena.release();
    return rc;
}

LVT_API int lvtk_detect(lvtk_ctx *c, const uint8_t *img, int rows, int cols, int stride, lvtk_keypoint *out, int cap,
                        int *n_out)
{
    if (!c || !img || !n_out || rows != c->params.img_height || cols != c->params.img_width || stride < cols)
        return LVTK_ERR_ARG;
    cudaSetDevice(c->device);
    if (int rc = ctx_upload_image(c, 0, img, rows, cols, stride))
        return rc;
    if (int rc = launch_detect(c->pool, c->ws, c->dp, c->d_slots, 1, c->feats_d, 0, 1, c->stream))
        return rc;
    if (int e = ctx_check_error(c))
        return e;
    return ctx_download_features(c, 0, out, nullptr, cap, n_out);
}

LVT_API int lvtk_brief(lvtk_ctx *c, const uint8_t *img, int rows, int cols, int stride, const lvtk_keypoint *in,
                       int n_in, lvtk_keypoint *out_kps, uint8_t *out_desc, int *n_out)
{
    if (!c || !img || !n_out || n_in < 0 || rows != c->params.img_height || cols != c->params.img_width || stride < cols)
        return LVTK_ERR_ARG;
    if (n_in > c->fcap)
        return LVTK_ERR_CAPACITY;
    cudaSetDevice(c->device);
    if (int rc = ctx_sync_pairs(c, c->stream))
        return rc;
    if (int rc = ctx_upload_image(c, 0, img, rows, cols, stride))
        return rc;
    std::vector<float2> xy(std::max(n_in, 1));
    std::vector<float> resp(std::max(n_in, 1));
    for (int i = 0; i < n_in; i++)
    {
        xy[i] = make_float2(in[i].x, in[i].y);
        resp[i] = in[i].response;
    }
    LVT_CUDA_TRY(cudaMemcpyAsync(c->d_in_xy, xy.data(), sizeof(float2) * std::max(n_in, 1), cudaMemcpyHostToDevice, c->stream));
    LVT_CUDA_TRY(cudaMemcpyAsync(c->d_in_resp, resp.data(), sizeof(float) * std::max(n_in, 1), cudaMemcpyHostToDevice, c->stream));
    LVT_CUDA_TRY(cudaMemcpyAsync(c->d_in_n, &n_in, sizeof(int), cudaMemcpyHostToDevice, c->stream));
    LVT_CUDA_TRY(cudaStreamSynchronize(c->stream));
    if (int rc = launch_border_filter(c->d_in_xy, c->d_in_resp, c->d_in_n, c->in_cap, c->feats_d, 1, rows, cols, c->ws.error,
                                      c->stream))
        return rc;
    if (int rc = launch_brief(c->pool, c->d_slots, 1, c->feats_d, c->d_brief_offsets, c->stream))
        return rc;
    if (int e = ctx_check_error(c))
        return e;
    return ctx_download_features(c, 0, out_kps, out_desc, n_in, n_out);
}

LVT_API int lvtk_extract(lvtk_ctx *c, const uint8_t *img, int rows, int cols, int stride, lvtk_keypoint *out_kps,
                         uint8_t *out_desc, int cap, int *n_out)
{
    if (!c || !img || !n_out || rows != c->params.img_height || cols != c->params.img_width || stride < cols)
        return LVTK_ERR_ARG;
    cudaSetDevice(c->device);
    if (int rc = ctx_sync_pairs(c, c->stream))
        return rc;
    if (int rc = ctx_upload_image(c, 0, img, rows, cols, stride))
        return rc;
    if (int rc = launch_detect(c->pool, c->ws, c->dp, c->d_slots, 1, c->feats_d, kBriefBorder, 1, c->stream))
        return rc;
    if (int rc = launch_brief(c->pool, c->d_slots, 1, c->feats_d, c->d_brief_offsets, c->stream))
        return rc;
    if (int e = ctx_check_error(c))
        return e;
    return ctx_download_features(c, 0, out_kps, out_desc, cap, n_out);
}

LVT_API int lvtk_match_projected(lvtk_ctx *c, const double *pts_xyz, const uint8_t *pts_desc, int m, const double q[4],
                                 const double t[3], const lvtk_keypoint *kps, const uint8_t *desc, int n,
                                 uint8_t *matched_flags, int retry_below, int *out_match_idx, float *out_d1,
                                 float *out_d2, int *out_count, int *retried)
{
    if (!c || m < 0 || n < 0 || !out_match_idx || !q || !t)
        return LVTK_ERR_ARG;
    if (m > std::min(c->in_cap, c->pcap) || n > c->fcap)
        return LVTK_ERR_CAPACITY;
    cudaSetDevice(c->device);
    if (int rc = ctx_upload_features(c, 0, kps, desc, n, matched_flags))
        return rc;
    if (m)
    {
        LVT_CUDA_TRY(cudaMemcpyAsync(c->map.xyz, pts_xyz, sizeof(double) * 3 * m, cudaMemcpyHostToDevice, c->stream));
        LVT_CUDA_TRY(cudaMemcpyAsync(c->map.desc, pts_desc, (size_t)32 * m, cudaMemcpyHostToDevice, c->stream));
    }
    if (int rc = launch_match_seam(c->map.xyz, c->map.desc, m, make_pose(q, t), c->feats_d, c->cam, retry_below, c->sc.ms,
                                   c->sc.map_cand, c->d_int_a, c->d_f_a, c->d_f_b, c->d_int_b, c->tcfg, c->stream))
        return rc;
    int cr[2] = {0, 0};
    if (m)
    {
        LVT_CUDA_TRY(cudaMemcpyAsync(out_match_idx, c->d_int_a, sizeof(int) * m, cudaMemcpyDeviceToHost, c->stream));
        if (out_d1)
            LVT_CUDA_TRY(cudaMemcpyAsync(out_d1, c->d_f_a, sizeof(float) * m, cudaMemcpyDeviceToHost, c->stream));
        if (out_d2)
            LVT_CUDA_TRY(cudaMemcpyAsync(out_d2, c->d_f_b, sizeof(float) * m, cudaMemcpyDeviceToHost, c->stream));
    }
    if (matched_flags && n)
        LVT_CUDA_TRY(cudaMemcpyAsync(matched_flags, c->feats_h[0].matched, n, cudaMemcpyDeviceToHost, c->stream));
    LVT_CUDA_TRY(cudaMemcpyAsync(cr, c->d_int_b, sizeof(cr), cudaMemcpyDeviceToHost, c->stream));
    LVT_CUDA_TRY(cudaStreamSynchronize(c->stream));
    if (out_count)
        *out_count = cr[0];
    if (retried)
        *retried = cr[1];
    return LVTK_OK;
}

LVT_API int lvtk_row_match(lvtk_ctx *c, const lvtk_keypoint *kl, const uint8_t *dl, int nl, uint8_t *ml,
                           const lvtk_keypoint *kr, const uint8_t *dr, int nr, uint8_t *mr, int *out_query,
                           int *out_train, int *n_matches)
{
    if (!c || nl < 0 || nr < 0 || !n_matches)
        return LVTK_ERR_ARG;
    if (nl > c->fcap || nr > c->fcap)
        return LVTK_ERR_CAPACITY;
    cudaSetDevice(c->device);
    if (int rc = ctx_upload_features(c, 0, kl, dl, nl, ml))
        return rc;
    if (int rc = ctx_upload_features(c, 1, kr, dr, nr, mr))
        return rc;
    if (int rc = launch_row_seam(c->feats_d, c->cam, c->row_cand[0], c->sc.row_choice, c->sc.ms.items, c->sc.pair_query,
                                 c->sc.pair_train, c->d_int_b, c->tcfg, c->stream))
        return rc;
    int np = 0;
    LVT_CUDA_TRY(cudaMemcpyAsync(&np, c->d_int_b, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
    LVT_CUDA_TRY(cudaStreamSynchronize(c->stream));
    *n_matches = np;
    if (np)
    {
        LVT_CUDA_TRY(cudaMemcpyAsync(out_query, c->sc.pair_query, sizeof(int) * np, cudaMemcpyDeviceToHost, c->stream));
        LVT_CUDA_TRY(cudaMemcpyAsync(out_train, c->sc.pair_train, sizeof(int) * np, cudaMemcpyDeviceToHost, c->stream));
    }
    if (ml && nl)
        LVT_CUDA_TRY(cudaMemcpyAsync(ml, c->feats_h[0].matched, nl, cudaMemcpyDeviceToHost, c->stream));
    if (mr && nr)
        LVT_CUDA_TRY(cudaMemcpyAsync(mr, c->feats_h[1].matched, nr, cudaMemcpyDeviceToHost, c->stream));
    LVT_CUDA_TRY(cudaStreamSynchronize(c->stream));
    return LVTK_OK;
}

LVT_API int lvtk_solve_pose(lvtk_ctx *c, const double *pts_xyz, const float *uv, int m, const double q_in[4],
                            const double t_in[3], double q_out[4], double t_out[3], uint8_t *inlier_marks)
{
    if (!c || m < 0 || !q_in || !t_in || !q_out || !t_out)
        return LVTK_ERR_ARG;
    if (m > std::min(c->in_cap, c->pcap))
        return LVTK_ERR_CAPACITY;
    cudaSetDevice(c->device);
    if (m)
    {
        LVT_CUDA_TRY(cudaMemcpyAsync(c->sc.sol_xyz, pts_xyz, sizeof(double) * 3 * m, cudaMemcpyHostToDevice, c->stream));
        LVT_CUDA_TRY(cudaMemcpyAsync(c->sc.sol_uv, uv, sizeof(float) * 2 * m, cudaMemcpyHostToDevice, c->stream));
    }
    if (int rc = launch_pose_seam(c->sc.sol_xyz, c->sc.sol_uv, m, make_pose(q_in, t_in), c->cam, c->sc.level, c->sc.inlier,
                                  c->sc.e2, c->d_pose_out, c->d_int_b, c->stream))
        return rc;
    PoseD out;
    LVT_CUDA_TRY(cudaMemcpyAsync(&out, c->d_pose_out, sizeof(out), cudaMemcpyDeviceToHost, c->stream));
    if (inlier_marks && m)
        LVT_CUDA_TRY(cudaMemcpyAsync(inlier_marks, c->sc.inlier, m, cudaMemcpyDeviceToHost, c->stream));
    LVT_CUDA_TRY(cudaStreamSynchronize(c->stream));
    q_out[0] = out.q.w;
    q_out[1] = out.q.x;
    q_out[2] = out.q.y;
    q_out[3] = out.q.z;
    t_out[0] = out.t[0];
    t_out[1] = out.t[1];
    t_out[2] = out.t[2];
    return LVTK_OK;
}

LVT_API int lvtk_rectify_maps(lvtk_ctx *c, const lvt_rectify_c *r, int rows, int cols, float *map_x, float *map_y)
{
    if (!c || !r || !map_x || !map_y || rows <= 0 || cols <= 0)
        return LVTK_ERR_ARG;
    LVT_CUDA_TRY(cudaSetDevice(c->device));
    RectifyDev dev;
    if (int rc = make_rectify_dev(*r, &dev))
        return rc;
    float *d_maps = nullptr;
    const size_t n = (size_t)rows * cols;
    LVT_CUDA_TRY(cudaMalloc(&d_maps, 2 * n * sizeof(float)));
    int rc = launch_rectify_maps(dev, rows, cols, d_maps, d_maps + n, c->stream);
    if (rc == LVTK_OK && (cudaMemcpyAsync(map_x, d_maps, n * sizeof(float), cudaMemcpyDeviceToHost, c->stream) != cudaSuccess ||
                          cudaMemcpyAsync(map_y, d_maps + n, n * sizeof(float), cudaMemcpyDeviceToHost, c->stream) != cudaSuccess ||
                          cudaStreamSynchronize(c->stream) != cudaSuccess))
        rc = LVTK_ERR_CUDA;
    cudaFree(d_maps);
    return rc;
}

LVT_API int lvtk_rectify(lvtk_ctx *c, const uint8_t *raw, int rows, int cols, int stride, const lvt_rectify_c *r,
                         uint8_t *out)
{
    if (!c || !raw || !r || !out || rows != c->params.img_height || cols != c->params.img_width || stride < cols)
        return LVTK_ERR_ARG;
    LVT_CUDA_TRY(cudaSetDevice(c->device));
    if (int rc = make_rectify_dev(*r, &c->rect[0]))
        return rc;
    ctx_stage_image(c, 0, raw, rows, cols, stride, true);
    if (int rc = ctx_stage_flush(c))
        return rc;
    if (int rc = ctx_rectify_slot(c, 0, 0, c->stream))
        return rc;
    LVT_CUDA_TRY(cudaMemcpy2DAsync(out, cols, c->pool.data, c->pool.pitch, cols, rows, cudaMemcpyDeviceToHost, c->stream));
    LVT_CUDA_TRY(cudaStreamSynchronize(c->stream));
    return LVTK_OK;
}

LVT_API int lvtk_triangulate(lvtk_ctx *c, const double q[4], const double t[3], const float *uv_left,
                             const float *uv_right, int n, double *out_xyz, uint8_t *out_valid)
{
    if (!c || n < 0 || !q || !t)
        return LVTK_ERR_ARG;
    if (n > std::min(c->in_cap, c->pcap))
        return LVTK_ERR_CAPACITY;
    if (n == 0)
        return LVTK_OK;
    cudaSetDevice(c->device);
    LVT_CUDA_TRY(cudaMemcpyAsync(c->sc.sol_uv, uv_left, sizeof(float) * 2 * n, cudaMemcpyHostToDevice, c->stream));
    LVT_CUDA_TRY(cudaMemcpyAsync(c->sc.ms.proj, uv_right, sizeof(float) * 2 * n, cudaMemcpyHostToDevice, c->stream));
    if (int rc = launch_tri_seam(make_pose(q, t), c->cam, c->sc.sol_uv, c->sc.ms.proj, n, c->sc.tri_xyz, c->sc.tri_ok,
                                 c->stream))
        return rc;
    LVT_CUDA_TRY(cudaMemcpyAsync(out_xyz, c->sc.tri_xyz, sizeof(double) * 3 * n, cudaMemcpyDeviceToHost, c->stream));
    LVT_CUDA_TRY(cudaMemcpyAsync(out_valid, c->sc.tri_ok, n, cudaMemcpyDeviceToHost, c->stream));
    LVT_CUDA_TRY(cudaStreamSynchronize(c->stream));
    return LVTK_OK;
}

} /* extern "C" */
