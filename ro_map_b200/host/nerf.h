// nerf.h — nerf::NeRF, one object: mirrors MON/Core/include/nerf.h:19-88 (same method names, argument meaning and
// threading contract) on top of the C ABI of the B200 core (include/mon_c.h).
#pragma once
#include <condition_variable>
#include <memory>
#include <string>
#include <vector>

#include "common.h"

struct mon_object;

namespace nerf {

class NeRF_Dataset;

class NeRF {
public:
    NeRF();
    ~NeRF();
    // offline
    bool CreateModelOffline(const std::string path, bool useDenseDepth);
    bool ReadBboxOffline(const std::string path);
    void TrainOffline(const int iterations);
    // online
    void SetAttributes(const int Class, const Eigen::Matrix4f& ObjTow, const BoundingBox& BoundingBox, size_t numBbox);
    bool CreateModelOnline(bool useSparseDepth, int Iterations);
    void TrainOnline();
    void UpdateFrameBBox(const std::vector<nerf::FrameIdAndBbox>& vFrameBbox, const int train_step);
    bool CheckFinish();
    void RequestFinish();
    // tools
    void RenderTestImg(const std::string out_path, const std::vector<std::string>& timestamp, const std::vector<Eigen::Matrix4f>& testTwc,
                       const std::vector<FrameIdAndBbox>& testBbox, const float radius);
    // NeRF_Model::RenderVideo (nerf_model.cu:1832-1991): 60 turn-table views into <folder>/<i>.png
    void RenderVideo(const std::string img_path_folder, const std::string depth_path_folder, const float radius);
    std::vector<Eigen::Matrix4f> GetTwc();
    BoundingBox GetBoundingBox();
    Eigen::Matrix4f GetObjTow();
    CPUMeshData& GetCPUMeshData();
    std::vector<FrameIdAndBbox> GetFrameIdAndBBox();
    void DrawCPUMesh();   // OpenGL client arrays where <GL/gl.h> exists at build time; a no-op in headless builds
    void DrawMesh();
    void SaveMesh(const std::string outname);   // NeRF_Model::SaveMesh: ASCII PLY of the current CPU mesh
    // additions (not in the reference): logged loss of the last Train_Step and its device time
    float LastLoss() const { return mfLastLoss; }
    float LastTrainMs() const { return mfLastMs; }
    int TrainingStep() const { return mnTrainingStep; }

    static int curId;
    int mId;
    static int GPUnum;
    static int curGPUid;
    int mGPUid = -1;
    int mClass = 0;
    uint8_t mInstanceId = 0;
    Eigen::Matrix4f mObjTow;
    BoundingBox mBoundingBox;
    std::vector<FrameIdAndBbox> mFrameIdBbox;
    size_t mnBbox = 0;
    std::mutex mUpdateBbox;
    std::condition_variable mCond;
    std::mutex mFinishMutex;
    bool mbFinishRequested = false;
    bool mbUseDepth = false;
    int mnIteration = 500;
    std::shared_ptr<NeRF_Dataset> mpTrainData;
    size_t mDataMutexIdx = 0;
    int mnTrainStep = 0;
    CPUMeshData mCPUMeshData;

private:
    bool CreateCore();
    void TrainStep(int iters);
    void UpdateMesh();
    mon_object* mpCore = nullptr;   // == NeRF_Model
    size_t mnCoreBbox = 0;          // boxes already uploaded (NeRF_Model::mnBbox)
    float mfLastLoss = 0.0f, mfLastMs = 0.0f;
    int mnTrainingStep = 0;
};

}  // namespace nerf
