// online_replay.cpp — replays the call sequence the ORB-SLAM2 frontend issues against nerf::NerfManagerOnline
// (src/System.cc:120-138, src/LocalMapping.cc:1122-1280) from an on-disk sequence, without the
// SLAM frontend (BASELINE.json config 5: the real 'scene1' recording is not available offline):
//
//   online_replay <network_config.json> <dataset_dir> <use_depth 0|1> [train_iters=500] [n_objects=4] [out_dir=output_online]
//
//   Init -> DatasetInit(fx,fy,cx,cy,H,W,n) -> per keyframe: NewFrameToDataset(id, stamp, bgr, instance, depth, Twc)
//   -> first observation of an object: CreateNeRF(class, Tow, bbox) -> every observation: UpdateNeRFBbox(idx, {box}, 1)
//   -> WaitThreadsEnd -> RenderNeRFsTest.
// Prints per-keyframe ingest latency and per-object totals.
#include <sys/stat.h>

#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdlib>
#include <fstream>
#include <iostream>
#include <map>
#include <sstream>
#include <vector>

#include "nerf_data.h"
#include "nerf_manager.h"
#include "png_io.h"

using clk = std::chrono::steady_clock;

struct ObjTrack {
    int cls = 0;
    Eigen::Matrix4f Tow;
    nerf::BoundingBox box;
    std::map<uint32_t, nerf::FrameIdAndBbox> obs;   // frame id -> 2-D box
    long nerf_idx = -1;
};

int main(int argc, char** argv) {
    if (argc < 4) {
        std::cerr << "usage: online_replay <network_config.json> <dataset_dir> <use_depth 0|1> [train_iters=500] [n_objects=4] [out_dir]" << std::endl;
        return 1;
    }
    const std::string cfg = argv[1], data = argv[2];
    const bool use_depth = std::atoi(argv[3]) != 0;
    const int iters = argc > 4 ? std::atoi(argv[4]) : 500;
    const int n_obj = argc > 5 ? std::atoi(argv[5]) : 4;
    const std::string out_dir = argc > 6 ? argv[6] : "output_online";

    // the sequence index (stamps, poses, intrinsics) through the same reader the offline path uses
    nerf::NeRF_Dataset index(use_depth);
    if (!index.ReadDataset(data)) return 2;
    std::vector<std::string> stamps(index.mnImages);
    for (auto& kv : index.mStampToIdx) stamps[kv.second] = kv.first;

    // object tracks from obj_offline/k.txt (class, object pose, half extents, per-frame 2-D boxes)
    std::vector<ObjTrack> tracks;
    for (int k = 0; k < n_obj; ++k) {
        std::ifstream f(data + "/obj_offline/" + std::to_string(k) + ".txt");
        if (!f) continue;
        std::string s;
        std::getline(f, s);
        std::getline(f, s);
        std::stringstream ss(s);
        ObjTrack t;
        float v[10];
        ss >> t.cls;
        for (float& x : v) ss >> x;
        t.Tow = mon_compat::pose_from_tq(v[0], v[1], v[2], v[3], v[4], v[5], v[6]).inverse();
        t.box.min = Eigen::Vector3f(-v[7], -v[8], -v[9]);
        t.box.max = Eigen::Vector3f(v[7], v[8], v[9]);
        while (std::getline(f, s)) {
            if (s.empty()) continue;
            std::stringstream ls(s);
            std::string stamp;
            nerf::FrameIdAndBbox b;
            ls >> stamp >> b.x >> b.y >> b.h >> b.w;
            b.FrameId = index.mStampToIdx[stamp];
            t.obs[b.FrameId] = b;
        }
        tracks.push_back(t);
    }
    if (tracks.empty()) { std::cerr << "no object tracks" << std::endl; return 2; }

    nerf::NerfManagerOnline manager(cfg, use_depth, iters);
    manager.Init();
    manager.DatasetInit(index.fx, index.fy, index.cx, index.cy, index.H, index.W, index.mnImages);

    const size_t px = (size_t)index.H * index.W;
    double ingest_ms = 0.0;
    std::vector<double> ingest_each;
    std::string err;
    for (uint32_t id = 0; id < index.mnImages; ++id) {
        png_io::Image rgb, inst, dep;
        if (!png_io::read(index.mvImagesPath[id], rgb, err) || !png_io::read(index.mvInstancesPath[id], inst, err)) { std::cerr << err << std::endl; return 3; }
        cv::Mat img(index.H, index.W, CV_8UC3), instance(index.H, index.W, CV_8UC1), depth;
        for (size_t p = 0; p < px; ++p) {   // the frontend hands over BGR (cv::imread order)
            uint8_t* d = img.ptr<uint8_t>(0, 0) + 3 * p;
            d[0] = rgb.u8[3 * p + 2]; d[1] = rgb.u8[3 * p + 1]; d[2] = rgb.u8[3 * p];
        }
        std::memcpy(instance.ptr<uint8_t>(0, 0), inst.u8.data(), px);
        if (use_depth) {
            if (!png_io::read(index.mvDepthsPath[id], dep, err)) { std::cerr << err << std::endl; return 3; }
            depth.create(index.H, index.W, CV_32FC1);
            float* dp = depth.ptr<float>(0, 0);
            for (size_t p = 0; p < px; ++p) dp[p] = (dep.bit_depth == 16 ? (float)dep.u16[p] : (float)dep.u8[p]) * index.mfDepthScale;
        }
        const auto t0 = clk::now();
        manager.NewFrameToDataset(id, stamps[id], img, instance, depth, index.mvIamgesPose[id]);
        ingest_each.push_back(std::chrono::duration<double, std::milli>(clk::now() - t0).count());
        ingest_ms += ingest_each.back();
        for (auto& t : tracks) {
            auto it = t.obs.find(id);
            if (it == t.obs.end()) continue;
            if (t.nerf_idx < 0) t.nerf_idx = (long)manager.CreateNeRF(t.cls, t.Tow, t.box);
            manager.UpdateNeRFBbox((size_t)t.nerf_idx, {it->second}, 1);
        }
    }
    manager.WaitThreadsEnd();
    std::cout << "ingest_ms_per_keyframe " << ingest_ms / index.mnImages << " keyframes " << index.mnImages << std::endl;
    {   // distribution of the frontend thread's NewFrameToDataset calls (objects training and meshing meanwhile)
        std::vector<double> v = ingest_each;
        std::sort(v.begin(), v.end());
        auto q = [&](double f) { return v.empty() ? 0.0 : v[std::min(v.size() - 1, (size_t)(f * (double)v.size()))]; };
        std::cout << "ingest_ms min " << q(0.0) << " median " << q(0.5) << " p90 " << q(0.9) << " max " << (v.empty() ? 0.0 : v.back()) << std::endl;
        std::cout << "ingest_ms_each";
        for (double x : ingest_each) std::cout << " " << std::round(x * 100.0) / 100.0;
        std::cout << std::endl;
    }
    mkdir(out_dir.c_str(), 0755);
    for (auto& t : tracks) {
        if (t.nerf_idx < 0) continue;
        auto obj = manager.mvpNeRFs[(size_t)t.nerf_idx];
        std::cout << "object " << obj->mId << " gpu " << obj->mGPUid << " boxes " << obj->mnBbox << " step " << obj->TrainingStep() << " loss "
                  << obj->LastLoss() << " ms_per_train_step " << obj->LastTrainMs() << std::endl;
        const nerf::FrameIdAndBbox first = t.obs.begin()->second;
        const float radius = 5.0f * std::max(t.box.max[0], std::max(t.box.max[1], t.box.max[2]));   // RenderRadius = mfMaxDist * 5 (src/System.cc:609)
        manager.RenderNeRFsTest(out_dir, (size_t)t.nerf_idx, {"view0"}, {first}, {index.mvIamgesPose[first.FrameId]}, radius);
    }
    return 0;
}
