// Compiled and run by tests/test_abi.py against the CPU oracle (the same C ABI as the CUDA library):
// the C++ mirror of lvt_system tracks a few synthetic frames read from stdin-free in-memory images.
#include "lvt_system.hpp"
#include <cstdio>
#include <cstdlib>
#include <vector>

int main(int argc, char **argv)
{
    if (argc != 3)
        return 2;
    const int W = std::atoi(argv[1]), H = std::atoi(argv[2]);
    lvt_params_c p;
    lvt_params_default(&p);
    p.fx = p.fy = 400.f;
    p.cx = W / 2.f;
    p.cy = H / 2.f;
    p.baseline = 0.5f;
    p.img_width = W;
    p.img_height = H;
    lvt_system *vo = lvt_system::create(p, lvt_system::eSensor_STEREO);
    if (!vo)
        return 3;
    if (vo->get_state() != lvt_system::eState_NOT_INITIALIZED || vo->get_sensor_type() != lvt_system::eSensor_STEREO)
        return 4;
    // a deterministic textured pair: the right image is the left one shifted by 8 pixels
    std::vector<unsigned char> L((size_t)W * H), R((size_t)W * H);
    unsigned s = 12345;
    std::vector<unsigned char> canvas((size_t)(W + 8) * H);
    for (size_t i = 0; i < canvas.size(); i++)
    {
        s = s * 1664525u + 1013904223u;
        canvas[i] = (unsigned char)((s >> 24) & 0xE0);
    }
    for (int y = 0; y < H; y++)
        for (int x = 0; x < W; x++)
        {
            L[(size_t)y * W + x] = canvas[(size_t)y * (W + 8) + x];
            R[(size_t)y * W + x] = canvas[(size_t)y * (W + 8) + x + 8];
        }
    const lvt_image_view l{L.data(), H, W}, r{R.data(), H, W};
    lvt_pose_view pose = vo->track(l, r);
    if (pose.R[0][0] != 1.0 || pose.t[0] != 0.0) // the first frame returns the identity (lvt_system.cpp:185-193)
        return 5;
    pose = vo->track(l, r); // same frame again: the camera has not moved
    const lvt_frame_info fi = vo->frame_info();
    std::printf("state %d tracked %d |t| %.3e\n", (int)vo->get_state(), fi.tracked, pose.t[0] * pose.t[0] + pose.t[1] * pose.t[1] + pose.t[2] * pose.t[2]);
    const bool ok = vo->get_state() == lvt_system::eState_TRACKING && fi.tracked > 50 &&
                    pose.t[0] * pose.t[0] + pose.t[1] * pose.t[1] + pose.t[2] * pose.t[2] < 1e-6;
    vo->reset();
    const bool reset_ok = vo->get_state() == lvt_system::eState_NOT_INITIALIZED;
    // the same two frames as a batch: same answer
    const std::vector<lvt_pose_view> batch = vo->track_batch(std::vector<lvt_image_view>(2, l), std::vector<lvt_image_view>(2, r));
    if (batch.size() != 2 || vo->last_status() != 0 || batch[0].R[0][0] != 1.0 || batch[1].t[0] != pose.t[0] ||
        vo->get_state() != lvt_system::eState_TRACKING)
        return 7;
    lvt_system::destroy(vo);
    return ok && reset_ok ? 0 : 6;
}
