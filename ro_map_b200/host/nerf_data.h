// nerf_data.h — nerf::NeRF_Dataset, the keyframe set resident on one GPU (MON/Core/include/nerf_data.h:19-72).
// Same members the managers and NeRF touch; the device side is a mon_dataset handle of the B200 core, which keeps
// u8 pixels (4x less HBM and PCIe traffic than the reference's float pixels) and converts in-kernel.
#pragma once
#include <map>
#include <memory>
#include <mutex>
#include <string>
#include <vector>

#include "common.h"

struct mon_dataset;

namespace nerf {

class NeRF_Dataset {
public:
    explicit NeRF_Dataset(bool useDepth) : mbUseDepth(useDepth) {}
    ~NeRF_Dataset();
    bool ReadDataset(const std::string datasetPath);   // config.yaml, img.txt, groundtruth.txt
    bool DataToGPU();                                  // offline: decode every frame and upload it
    bool InitDataToGPU();                              // online: allocate for mnImages frames
    void FrameDataToGPU(unsigned int imgId, const std::string timestamp);
    void UpdateDataGPU(unsigned int CurId, unsigned int FrameNum);

    int mGPUid = 0;
    bool mbUseDepth = false;
    float fx = 0, fy = 0, cx = 0, cy = 0;
    int H = 0, W = 0;
    float mfDepthScale = 1.0f;
    // offline: the depth PNGs' 16-bit samples go to the GPU as they are and the batch kernel applies mfDepthScale to the pixels it
    // picks (mon_dataset_set_depth_u16) instead of a host-side convertTo pass and a float plane (nerf_data.cu:176-186)
    bool mbDepthRaw16 = false;
    size_t mnImages = 0;
    size_t mFrameDataNum = 0;
    std::vector<std::string> mvImagesPath, mvDepthsPath, mvInstancesPath;
    std::map<std::string, uint32_t> mStampToIdx;
    std::vector<Eigen::Matrix4f> mvIamgesPose;   // (sic)
    // online hand-over slots, as in the reference
    cv::Mat Temp_Img, Temp_Instance, Temp_Depth;
    Eigen::Matrix4f Temp_Pose;
    std::vector<Eigen::Matrix4f> mvTemp_Update_Pose;
    std::vector<std::unique_ptr<std::mutex>> mvUpdateMutex;

    mon_dataset* mpCore = nullptr;
};

}  // namespace nerf
