/*
 * CPU ORACLE (test infrastructure) -- BRIEF-32 descriptors.
 *
 * Restates cv::xfeatures2d::BriefDescriptorExtractor::create()->compute() (32 bytes, no
 * orientation), the call at lvt/src/lvt_image_features_handler.cpp:172 (also :190, :247).
 * opencv_contrib (modules/xfeatures2d/src/brief.cpp, generated_32.i; same release as the
 * OpenCV 3.x the reference links, README.md:12) is not in /root/reference and not in this
 * image, so its published algorithm is restated:
 *   1. sum = integral(image)                          int32, (H+1) x (W+1)
 *   2. KeyPointsFilter::runByImageBorder(kps, size, 28)   28 = PATCH_SIZE/2 + KERNEL_SIZE/2
 *   3. byte b, bit (7-j) = smoothedSum(y1,x1) < smoothedSum(y2,x2) for test 8b+j,
 *      smoothedSum = 9x9 box sum centred at ((int)(pt.x+.5)+x, (int)(pt.y+.5)+y)
 * The 256 test pairs are an INPUT: brief_pairs.inc is a stand-in table
 * (tools/gen_brief_pairs.py); lvt_set_brief_pairs() injects the genuine one.
 * PARITY: bits are exact against this restatement, unpinned against OpenCV.
 */
#include "lvto.h"
#include <algorithm>

namespace lvto
{

static const signed char DEFAULT_PAIRS[256][4] = {
#include "brief_pairs.inc"
};
static signed char g_pairs[256][4];
static bool g_pairs_init = false;

void brief_set_pairs(const signed char pairs[256][4])
{
    std::memcpy(g_pairs, pairs ? pairs : DEFAULT_PAIRS, sizeof(g_pairs));
    g_pairs_init = true;
}

void brief_compute(const Image &img, std::vector<Keypoint> &kps, std::vector<Desc> &desc)
{
    if (!g_pairs_init)
        brief_set_pairs(nullptr);
    const int W = img.cols, H = img.rows;
    const int border = 28;

    /* runByImageBorder: Rect(Point(b,b), Point(W-b,H-b)).contains(Point(pt)), where the
     * Point2f -> Point2i conversion is saturate_cast<int> == round-half-to-even */
    std::vector<Keypoint> kept;
    if (!(H <= border * 2 || W <= border * 2))
    {
        kept.reserve(kps.size());
        for (const Keypoint &k : kps)
        {
            const long rx = lrintf(k.x), ry = lrintf(k.y);
            if (rx >= border && rx < W - border && ry >= border && ry < H - border)
                kept.push_back(k);
        }
    }
    kps.swap(kept);

    desc.assign(kps.size(), Desc{});
    if (kps.empty())
        return;

    std::vector<int32_t> sum((size_t)(H + 1) * (W + 1), 0);
    const int sw = W + 1;
    for (int y = 0; y < H; y++)
    {
        int32_t rowacc = 0;
        for (int x = 0; x < W; x++)
        {
            rowacc += img.at(y, x);
            sum[(size_t)(y + 1) * sw + (x + 1)] = sum[(size_t)y * sw + (x + 1)] + rowacc;
        }
    }
    auto smoothed = [&](int X, int Y, int dy, int dx) -> int32_t {
        const int iy = Y + dy, ix = X + dx;
        return sum[(size_t)(iy + 5) * sw + (ix + 5)] - sum[(size_t)(iy + 5) * sw + (ix - 4)] -
               sum[(size_t)(iy - 4) * sw + (ix + 5)] + sum[(size_t)(iy - 4) * sw + (ix - 4)];
    };
    for (size_t i = 0; i < kps.size(); i++)
    {
        int X = (int)(kps[i].x + 0.5), Y = (int)(kps[i].y + 0.5); /* double add, as OpenCV */
        /* a fractional external corner at exactly k+0.5 (k even) passes the round-half-even border
         * filter yet samples one pixel further; OpenCV then reads past the integral image (undefined).
         * Defined here, and in the CUDA path, as clamping to the last valid centre. */
        X = std::min(X, W - border - 1);
        Y = std::min(Y, H - border - 1);
        for (int b = 0; b < 32; b++)
        {
            unsigned v = 0;
            for (int j = 0; j < 8; j++)
            {
                const signed char *t = g_pairs[8 * b + j];
                v |= (unsigned)(smoothed(X, Y, t[0], t[1]) < smoothed(X, Y, t[2], t[3])) << (7 - j);
            }
            desc[i].b[b] = (uint8_t)v;
        }
    }
}

} // namespace lvto
