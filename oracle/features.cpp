/*
 * CPU ORACLE (test infrastructure) -- per-image feature container and masked matcher.
 *
 * Restates lvt/src/lvt_image_features_struct.cpp:35-148 (+ .h:82-85) and the OpenCV call it
 * makes, cv::BFMatcher(NORM_HAMMING,false).knnMatch(1xD, NxD, k=2, mask) (:104, :140), whose
 * observable behaviour (best two by (distance, index), fewer than two
 * results when fewer than two candidates) is pinned against cv2 4.13 in
 * tests/test_oracle_cv2.py.
 *
 * Built with -ffp-contract=off: dx*dx + dy*dy is two roundings + one, as in plain IEEE fp32.
 */
#include "lvto.h"
#include <algorithm>

namespace lvto
{

int hamming256(const uint8_t *a, const uint8_t *b)
{
    int d = 0;
    for (int i = 0; i < 32; i += 8)
    {
        uint64_t x, y;
        std::memcpy(&x, a + i, 8);
        std::memcpy(&y, b + i, 8);
        d += __builtin_popcountll(x ^ y);
    }
    return d;
}

/* best two of the candidates by (hamming, index) -- knnMatch(k=2) with a mask */
static inline void top2_push(Top2 &t, int idx, int dist)
{
    /* candidates arrive in arbitrary index order: compare (dist, idx) lexicographically */
    auto less = [](int d0, int i0, int d1, int i1) { return d0 < d1 || (d0 == d1 && i0 < i1); };
    if (t.n == 0)
    {
        t.idx[0] = idx;
        t.dist[0] = dist;
        t.n = 1;
    }
    else if (less(dist, idx, t.dist[0], t.idx[0]))
    {
        t.idx[1] = t.idx[0];
        t.dist[1] = t.dist[0];
        t.idx[0] = idx;
        t.dist[0] = dist;
        t.n = 2;
    }
    else if (t.n == 1 || less(dist, idx, t.dist[1], t.idx[1]))
    {
        t.idx[1] = idx;
        t.dist[1] = dist;
        t.n = 2;
    }
}

/* lvt/src/lvt_image_features_struct.cpp:35-66 */
void FeatureSet::init(int in_rows, int in_cols, std::vector<Keypoint> &in_kps, std::vector<Desc> &in_desc,
                      const lvt_params_c &p, const std::vector<float> *kps_depth)
{
    cell_size = 25;              /* LVT_HASHING_CELL_SIZE, lvt/src/lvt_definitions.h:32 */
    vertical_search_radius = 2;  /* LVT_ROW_MATCHING_VERTICAL_SEARCH_RADIUS, :31 */
    triangulation_ratio_th = p.triangulation_ratio_test_threshold;
    tracking_ratio_th = p.tracking_ratio_test_threshold;
    desc_dist_th = p.descriptor_matching_threshold;
    rows = in_rows;
    cols = in_cols;
    tracking_radius = p.tracking_radius;
    const float k_cell = (float)cell_size;
    cells_x = (int)std::ceil(cols / k_cell);
    cells_y = (int)std::ceil(rows / k_cell);
    kps.swap(in_kps);
    desc.swap(in_desc);
    cell_search_radius = (tracking_radius == cell_size) ? 1 : (int)std::ceil((float)tracking_radius / k_cell);
    grid.assign((size_t)cells_x * cells_y, std::vector<int>());
    for (int i = 0, n = (int)kps.size(); i < n; i++)
    {
        int hy = (int)std::floor(kps[i].y / k_cell), hx = (int)std::floor(kps[i].x / k_cell);
        /* RGB-D keypoints are undistorted after description and may hash outside the grid
         * (undefined behaviour in the reference): clamp. */
        hy = std::min(std::max(hy, 0), cells_y - 1);
        hx = std::min(std::max(hx, 0), cells_x - 1);
        grid[(size_t)hy * cells_x + hx].push_back(i);
    }
    reset_matched();
    depths.clear();
    if (kps_depth)
        depths = *kps_depth;
}

/* lvt/src/lvt_image_features_struct.cpp:68-120 */
int FeatureSet::find_match_index(double px, double py, const Desc &d, float *d1, float *d2) const
{
    const float ptx = (float)px, pty = (float)py;
    const float k_cell = (float)cell_size;
    const int hy = (int)std::floor(pty / k_cell), hx = (int)std::floor(ptx / k_cell);
    int start_y = std::max(hy - cell_search_radius, 0);
    int end_y = std::min(hy + cell_search_radius + 1, cells_y);
    int start_x = std::max(hx - cell_search_radius, 0);
    int end_x = std::min(hx + cell_search_radius + 1, cells_x);

    const float r2 = (float)(tracking_radius * tracking_radius);
    Top2 best;
    for (int i = start_y; i < end_y; i++)
    {
        for (int k = start_x; k < end_x; k++)
        {
            for (int kp_idx : grid[(size_t)i * cells_x + k])
            {
                if (matched[kp_idx])
                    continue;
                const float dx = kps[kp_idx].x - ptx;
                const float dy = kps[kp_idx].y - pty;
                if ((dx * dx + dy * dy) < r2)
                    top2_push(best, kp_idx, hamming256(d.b, desc[kp_idx].b));
            }
        }
    }

    if (best.n > 1)
    {
        const float d_ratio = (float)best.dist[0] / (float)best.dist[1];
        if (d_ratio < tracking_ratio_th)
        {
            *d1 = (float)best.dist[0];
            *d2 = (float)best.dist[1];
            return best.idx[0];
        }
    }
    else if (best.n == 1 && (float)best.dist[0] <= desc_dist_th)
    {
        *d1 = (float)best.dist[0];
        *d2 = -1.0f;
        return best.idx[0];
    }
    return -1;
}

/* lvt/src/lvt_image_features_struct.cpp:122-148 */
int FeatureSet::row_match(float ptx, float pty, const Desc &d) const
{
    (void)ptx; /* the reference applies no x / disparity constraint */
    int start_y = std::max((int)pty - vertical_search_radius, 0);
    int end_y = std::min((int)pty + vertical_search_radius, rows);

    Top2 best;
    for (int i = 0, n = (int)kps.size(); i < n; i++)
    {
        if (!matched[i] && kps[i].y >= (float)start_y && kps[i].y <= (float)end_y)
            top2_push(best, i, hamming256(d.b, desc[i].b));
    }
    if ((best.n > 1 && ((float)best.dist[0] / (float)best.dist[1]) < triangulation_ratio_th) ||
        (best.n == 1 && (float)best.dist[0] <= desc_dist_th))
        return best.idx[0];
    return -1;
}

/* lvt/src/lvt_image_features_handler.cpp:302-323 */
void row_match_all(FeatureSet *left, FeatureSet *right, std::vector<int> *query, std::vector<int> *train)
{
    for (int i = 0, n = left->size(); i < n; i++)
    {
        if (left->matched[i])
            continue;
        const int m = right->row_match(left->kps[i].x, left->kps[i].y, left->desc[i]);
        if (m != -1)
        {
            query->push_back(i);
            train->push_back(m);
            left->matched[i] = 1;
            right->matched[m] = 1;
        }
    }
}

} // namespace lvto
