/*
 * CPU ORACLE (test infrastructure) -- parameters.
 * Restates lvt/src/lvt_parameters.cpp:29-52 (defaults) and :54-93 (cv::FileStorage YAML:
 * flat "key: value" lines under a "%YAML:1.0" header; a missing key reads as 0).
 */
#include "lvto.h"
#include <cstdio>
#include <cstdlib>
#include <map>
#include <string>

namespace lvto
{

void params_default(lvt_params_c *p)
{
    std::memset(p, 0, sizeof(*p));
    p->fx = p->fy = p->cx = p->cy = 0.5f;
    p->img_width = p->img_height = 0;
    p->baseline = 0.f;
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

int params_from_file(lvt_params_c *p, const char *file)
{
    FILE *f = file ? std::fopen(file, "r") : nullptr;
    if (!f)
        return 0;
    std::map<std::string, double> kv;
    char line[1024];
    while (std::fgets(line, sizeof(line), f))
    {
        std::string s(line);
        const size_t hash = s.find('#');
        if (hash != std::string::npos)
            s.erase(hash);
        if (s.empty() || s[0] == '%' || s[0] == '-')
            continue;
        const size_t colon = s.find(':');
        if (colon == std::string::npos)
            continue;
        std::string key = s.substr(0, colon), val = s.substr(colon + 1);
        auto trim = [](std::string &t) {
            const char *ws = " \t\r\n\"";
            const size_t b = t.find_first_not_of(ws);
            if (b == std::string::npos)
            {
                t.clear();
                return;
            }
            t = t.substr(b, t.find_last_not_of(ws) - b + 1);
        };
        trim(key);
        trim(val);
        if (key.empty() || val.empty())
            continue;
        char *end = nullptr;
        const double v = std::strtod(val.c_str(), &end);
        if (end == val.c_str())
            continue;
        kv[key] = v;
    }
    std::fclose(f);
    auto get = [&kv](const char *k) -> double {
        auto it = kv.find(k);
        return it == kv.end() ? 0.0 : it->second;
    };
    auto geti = [&get](const char *k) -> int { return (int)std::lrint(get(k)); };
    std::memset(p, 0, sizeof(*p));
    p->fx = (float)get("fx");
    p->fy = (float)get("fy");
    p->cx = (float)get("cx");
    p->cy = (float)get("cy");
    p->k1 = (float)get("k1");
    p->k2 = (float)get("k2");
    p->p1 = (float)get("p1");
    p->p2 = (float)get("p2");
    p->k3 = (float)get("k3");
    p->baseline = (float)get("baseline");
    p->img_width = geti("img_width");
    p->img_height = geti("img_height");
    p->near_plane_distance = (float)get("near_plane_distance");
    p->far_plane_distance = (float)get("far_plane_distance");
    p->triangulation_ratio_test_threshold = (float)get("triangulation_ratio_test_threshold");
    p->tracking_ratio_test_threshold = (float)get("tracking_ratio_test_threshold");
    p->min_num_matches_for_tracking = geti("min_num_matches_for_tracking");
    p->tracking_radius = geti("tracking_radius");
    p->agast_threshold = geti("agast_threshold");
    p->untracked_threshold = geti("untracked_threshold");
    p->staged_threshold = geti("staged_threshold");
    p->descriptor_matching_threshold = (float)get("descriptor_matching_threshold");
    p->detection_cell_size = geti("detection_cell_size");
    p->max_keypoints_per_cell = geti("max_keypoints_per_cell");
    p->enable_logging = geti("enable_logging") != 0;
    p->enable_visualization = geti("enable_visualization") != 0;
    p->triangulation_policy = geti("triangulation_policy");
    p->viewer_camera_size = (float)get("viewer_camera_size");
    p->viewer_point_size = geti("viewer_point_size");
    return 1;
}

} // namespace lvto
