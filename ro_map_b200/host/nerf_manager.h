// nerf_manager.h — nerf::NerfManagerOffline / nerf::NerfManagerOnline, the interface MON/main.cpp and the
// ORB-SLAM2 frontend (src/System.cc:120-138, src/LocalMapping.cc:1122-1280, src/MapDrawer.cc:396) program against.
// Mirrors MON/Core/include/nerf_manager.h:21-91: same constructors, methods, public members and threading
// (one std::thread per object, objects placed round-robin on the visible GPUs, one dataset replica per GPU).
#pragma once
#include <iostream>
#include <memory>
#include <string>
#include <thread>
#include <vector>

#include "nerf.h"

namespace nerf {

class NeRF_Dataset;

class NerfManagerOffline {
public:
    NerfManagerOffline(const std::string datasetPath, const std::string networkConfigFile, bool useDenseDepth);
    bool Init();
    bool ReadDataset();
    bool CreateNeRF(const std::string objectFile);
    bool WaitThreadsEnd();
    std::shared_ptr<NeRF> GetNeRF(int idx);
    std::vector<std::shared_ptr<NeRF>> GetAllNeRF();
    std::vector<Eigen::Matrix4f> GetAllTwc();
    void GetIntrinsics(float& fx, float& fy, float& cx, float& cy);

    std::string msNetworkConfigFile;
    std::string msDatasetPath;
    bool mbUseDenseDepth;
    int mNumGPU = 0;
    int mnTrainSteps = 10;   // addition: the reference hard-codes TrainOffline(10) (nerf_manager.cu:89)
    std::vector<std::shared_ptr<NeRF_Dataset>> mvpDataset;
    std::vector<std::shared_ptr<NeRF>> mvpNeRFs;
    std::vector<std::thread> mvThreads;
};

class NerfManagerOnline {
public:
    NerfManagerOnline(const std::string network_config_file, bool UseSparseDepth, int TrainStepIterations);
    bool Init();
    void DatasetInit(float fx, float fy, float cx, float cy, int H, int W, size_t imgs);
    void NewFrameToDataset(unsigned int imgId, const std::string timestamp, cv::Mat& img, cv::Mat& instance, const cv::Mat& depth_img,
                           const Eigen::Matrix4f& pose);
    void UpdateDataset(unsigned int CurId, unsigned int FrameNum, const std::vector<Eigen::Matrix4f>& Poses);
    size_t CreateNeRF(const int Class, const Eigen::Matrix4f& ObjTow, const nerf::BoundingBox& BoundingBox);
    int GetFrameIdx(double timastamp);
    void UpdateNeRFBbox(const size_t idx, const std::vector<nerf::FrameIdAndBbox>& vFrameBbox, const int train_step);
    void DrawMesh(size_t idx);
    bool WaitThreadsEnd();
    void RenderNeRFsTest(const std::string out_path, const size_t Idx, const std::vector<std::string>& timestamp,
                         const std::vector<FrameIdAndBbox>& vBbox, const std::vector<Eigen::Matrix4f>& vTwc, const float radius);

    std::string mNetworkConfigFile;
    bool mbUseSparseDepth;
    int mnTrainStepIterations;
    int mNumGPU = 0;
    std::vector<std::shared_ptr<NeRF_Dataset>> mvpDataset;
    std::vector<std::shared_ptr<NeRF>> mvpNeRFs;
    std::vector<std::thread> mvThreads;
};

}  // namespace nerf
