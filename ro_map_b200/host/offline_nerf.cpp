// offline_nerf.cpp — headless OfflineNeRF: MON/main.cpp:287-345 without the Pangolin viewer.
//   offline_nerf <network_config.json> <dataset_dir> <use_depth 0|1> [train_steps=10] [n_objects=4]
// Loads <dataset>/obj_offline/{0..n-1}.txt like the reference (main.cpp:307-319), trains every object on its own
// host thread (objects round-robin over the visible GPUs), renders the first test box of each object and prints one
// line per object: id, GPU, final loss, ms per 500-iteration Train_Step.
#include <algorithm>
#include <sys/stat.h>

#include <cstdlib>
#include <fstream>
#include <iostream>

#include "nerf_manager.h"

int main(int argc, char** argv) {
    if (argc < 4) {
        std::cerr << "usage: offline_nerf <network_config.json> <dataset_dir> <use_depth 0|1> [train_steps=10] [n_objects=4]" << std::endl;
        return 1;
    }
    const std::string cfg = argv[1], data = argv[2];
    const bool use_depth = std::atoi(argv[3]) != 0;
    const int steps = argc > 4 ? std::atoi(argv[4]) : 10;
    const int n_obj = argc > 5 ? std::atoi(argv[5]) : 4;

    nerf::NerfManagerOffline manager(data, cfg, use_depth);
    manager.mnTrainSteps = steps;
    manager.Init();
    manager.ReadDataset();
    int created = 0;
    for (int k = 0; k < n_obj; ++k) {
        const std::string objfile = data + "/obj_offline/" + std::to_string(k) + ".txt";
        if (!std::ifstream(objfile)) continue;
        if (manager.CreateNeRF(objfile)) ++created;
    }
    if (created == 0) {
        std::cerr << "no object files under " << data << "/obj_offline" << std::endl;
        return 2;
    }
    manager.WaitThreadsEnd();
    mkdir("output", 0755);
    for (auto& obj : manager.GetAllNeRF()) {
        std::cout << "object " << obj->mId << " gpu " << obj->mGPUid << " step " << obj->TrainingStep() << " loss " << obj->LastLoss()
                  << " ms_per_train_step " << obj->LastTrainMs() << std::endl;
        auto boxes = obj->GetFrameIdAndBBox();
        if (!boxes.empty()) {
            auto Twc = obj->GetTwc();
            const nerf::BoundingBox bb = obj->GetBoundingBox();
            const float radius = 5.0f * std::max(bb.max[0], std::max(bb.max[1], bb.max[2]));   // RenderRadius = mfMaxDist * 5 (src/System.cc:609)
            obj->RenderTestImg("output", {"view0"}, {Twc[0]}, {boxes[0]}, radius);
        }
    }
    return 0;
}
