/*
 * CPU ORACLE (test infrastructure) -- corner detection.
 *
 * Restates, as observable behaviour, the OpenCV calls the reference makes at
 * lvt/src/lvt_image_features_handler.cpp:116,139 (cv::AgastFeatureDetector, OAST_9_16, NMS on)
 * and restates the reference's own code at :34-83 (ANMS), :95-114 (tile grid),
 * :131-154 (per-tile detect) and :161-169 (low-corner retry).
 *
 * OpenCV's agast.cpp / agast_score.cpp are not in /root/reference; the behaviour below
 *  is pinned bit-exactly against cv2 4.13 by tests/test_oracle_cv2.py
 * and tests/golden/agast_*.npz.
 */
#include "lvto.h"
#include <algorithm>
#include <limits>

namespace lvto
{

/* the 16-pixel Bresenham ring of radius 3, in OpenCV's OAST_9_16 order */
static const int RING_DX[16] = {0, 1, 2, 3, 3, 3, 2, 1, 0, -1, -2, -3, -3, -3, -2, -1};
static const int RING_DY[16] = {3, 3, 2, 1, 0, -1, -2, -3, -3, -3, -2, -1, 0, 1, 2, 3};

/* Largest threshold t for which the pixel is still a 9-of-16 segment-test corner:
 * max over the 16 arcs of min(|difference|) over the arc, minus one (strict comparison).
 * This equals AGAST's binary-search score (agast_score.cpp) because the OAST_9_16 decision
 * tree accepts exactly the segment-test corners. */
static inline int corner_score(const int d[16])
{
    int best = -100000;
    for (int k = 0; k < 16; k++)
    {
        int mn = d[k], mx = d[k];
        for (int j = 1; j < 9; j++)
        {
            const int v = d[(k + j) & 15];
            mn = std::min(mn, v);
            mx = std::max(mx, v);
        }
        best = std::max(best, std::max(mn, -mx));
    }
    return best - 1;
}

static void agast_raw(const Image &img, int threshold, std::vector<Keypoint> &out)
{
    out.clear();
    int off[16];
    for (int i = 0; i < 16; i++)
        off[i] = RING_DY[i] * img.stride + RING_DX[i];
    /* x in [3, cols-4], y in [3, rows-4] -- OpenCV's OAST_9_16 scan bounds */
    for (int y = 3; y <= img.rows - 4; y++)
    {
        const uint8_t *row = img.data + (size_t)y * img.stride;
        for (int x = 3; x <= img.cols - 4; x++)
        {
            const uint8_t *p = row + x;
            const int c = *p;
            /* any 9-arc holds at least two (adjacent) compass points: cheap rejection */
            const int hi = c + threshold, lo = c - threshold;
            int nb = (p[off[0]] > hi) + (p[off[4]] > hi) + (p[off[8]] > hi) + (p[off[12]] > hi);
            int nd = (p[off[0]] < lo) + (p[off[4]] < lo) + (p[off[8]] < lo) + (p[off[12]] < lo);
            if (nb < 2 && nd < 2)
                continue;
            int d[16];
            for (int i = 0; i < 16; i++)
                d[i] = (int)p[off[i]] - c;
            const int s = corner_score(d);
            if (s >= threshold)
                out.push_back(Keypoint{(float)x, (float)y, (float)s});
        }
    }
}

/* OpenCV AGAST non-maximum suppression: one raster pass with a parent array; the
 * survivor of every 4-connected component is its maximum, ties broken by merge order. */
static void agast_nms(const std::vector<Keypoint> &k, std::vector<Keypoint> &out)
{
    const int n = (int)k.size();
    std::vector<int> flag(n, -1);
    int lastRow = 0, nextRow = 0, lastIdx = 0, nextIdx = 0;
    auto root = [&flag](int i) {
        while (flag[i] != -1)
            i = flag[i];
        return i;
    };
    for (int c = 0; c < n; c++)
    {
        const int cy = (int)k[c].y, cx = (int)k[c].x;
        if (lastRow + 1 < cy)
        {
            lastRow = nextRow;
            lastIdx = nextIdx;
        }
        if (nextRow != cy)
        {
            nextRow = cy;
            nextIdx = c;
        }
        if (lastRow + 1 == cy)
        {
            while ((int)k[lastIdx].x < cx && (int)k[lastIdx].y == lastRow)
                lastIdx++;
            if ((int)k[lastIdx].x == cx && lastIdx != c)
            {
                const int w = root(lastIdx);
                if (k[c].response < k[w].response)
                    flag[c] = w;
                else
                    flag[w] = c;
            }
        }
        int t = c - 1;
        if (c != 0 && (int)k[t].y == cy && (int)k[t].x + 1 == cx)
        {
            const int a = flag[c];
            t = root(t);
            if (a == -1)
            {
                if (t != c)
                {
                    if (k[c].response < k[t].response)
                        flag[c] = t;
                    else
                        flag[t] = c;
                }
            }
            else if (t != a)
            {
                if (k[a].response < k[t].response)
                {
                    flag[a] = t;
                    flag[c] = t;
                }
                else
                {
                    flag[t] = a;
                    flag[c] = a;
                }
            }
        }
    }
    out.clear();
    for (int i = 0; i < n; i++)
        if (flag[i] == -1)
            out.push_back(k[i]);
}

void agast_detect(const Image &img, int threshold, bool nonmax, std::vector<Keypoint> &out)
{
    std::vector<Keypoint> raw;
    agast_raw(img, threshold, raw);
    if (!nonmax)
    {
        out.swap(raw);
        return;
    }
    agast_nms(raw, out);
}

/* lvt/src/lvt_image_features_handler.cpp:34-83, statement for statement */
void anms(std::vector<Keypoint> &keypoints, int num_to_keep, float tx, float ty)
{
    std::sort(keypoints.begin(), keypoints.end(),
              [](const Keypoint &lhs, const Keypoint &rhs) { return lhs.response > rhs.response; });

    std::vector<Keypoint> kept;
    kept.reserve(num_to_keep);
    std::vector<float> radii(keypoints.size()), radii_sorted(keypoints.size());

    const float robust_coeff = 1.11;
    for (int i = 0, n = (int)keypoints.size(); i < n; i++)
    {
        const float response = keypoints[i].response * robust_coeff;
        float radius = (std::numeric_limits<float>::max)();
        for (int j = 0; j < i && keypoints[j].response > response; j++)
        {
            const float dx = keypoints[i].x - keypoints[j].x;
            const float dy = keypoints[i].y - keypoints[j].y;
            radius = (std::min)(radius, dx * dx + dy * dy);
        }
        radius = sqrtf(radius);
        radii[i] = radius;
        radii_sorted[i] = radius;
    }

    std::sort(radii_sorted.begin(), radii_sorted.end(), [](const float &l, const float &r) { return l > r; });

    const float decision_radius = radii_sorted[num_to_keep];
    for (int i = 0, n = (int)radii.size(); i < n; i++)
    {
        if (radii[i] >= decision_radius)
        {
            keypoints[i].x += tx;
            keypoints[i].y += ty;
            kept.push_back(keypoints[i]);
        }
    }
    kept.swap(keypoints);
}

/* lvt/src/lvt_image_features_handler.cpp:95-114 */
std::vector<Rect> tile_rects(int img_w, int img_h, int s)
{
    std::vector<Rect> rects;
    const int ny = 1 + ((img_h - 1) / s);
    const int nx = 1 + ((img_w - 1) / s);
    for (int i = 0; i < ny; i++)
    {
        for (int k = 0; k < nx; k++)
        {
            int sy = s;
            if ((i == ny - 1) && ((i + 1) * s > img_h))
                sy = img_h - (i * s);
            int sx = s;
            if ((k == nx - 1) && ((k + 1) * s > img_w))
                sx = img_w - (k * s);
            rects.push_back(Rect{k * s, i * s, sx, sy});
        }
    }
    return rects;
}

/* lvt/src/lvt_image_features_handler.cpp:131-154 */
static void detect_pass(const Image &img, const std::vector<Rect> &rects, int threshold, int max_per_cell,
                        std::vector<Keypoint> &all)
{
    for (const Rect &r : rects)
    {
        Image sub{img.data + (size_t)r.y * img.stride + r.x, r.h, r.w, img.stride};
        std::vector<Keypoint> kps;
        agast_detect(sub, threshold, true, kps);
        if ((long)kps.size() > (long)max_per_cell)
        {
            anms(kps, max_per_cell, (float)r.x, (float)r.y);
        }
        else
        {
            for (Keypoint &k : kps)
            {
                k.x += (float)r.x;
                k.y += (float)r.y;
            }
        }
        all.insert(all.end(), kps.begin(), kps.end());
    }
}

/* lvt/src/lvt_image_features_handler.cpp:158-169 */
void detect_corners(const Image &img, const lvt_params_c &p, std::vector<Keypoint> &out)
{
    /* the reference builds the rects from the configured image size (:95-114), not the frame's */
    const std::vector<Rect> rects = tile_rects(p.img_width, p.img_height, p.detection_cell_size);
    out.clear();
    detect_pass(img, rects, p.agast_threshold, p.max_keypoints_per_cell, out);
    if (out.size() < 200 /* LVT_CORNERS_LOW_TH, lvt/src/lvt_definitions.h:33 */)
    {
        out.clear();
        const int lowered = (int)((double)p.agast_threshold * 0.5 + 0.5);
        detect_pass(img, rects, lowered, p.max_keypoints_per_cell, out);
    }
}

} // namespace lvto
