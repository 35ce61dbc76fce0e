// common.h — POD types of the libMON.so boundary, mirroring MON/Core/include/common.h:18-54.
// (MeshData, the OpenGL VBO variant, is omitted: the reference itself only uses the CPU mesh path, nerf.cu:140-145.)
#pragma once
#include <cstdint>
#include <memory>
#include <mutex>
#include <string>
#include <vector>

#include "mon_compat.h"

namespace nerf {

struct FrameIdAndBbox {     // == mon_bbox2d; file order of obj_offline/k.txt is x, y, h, w
    uint32_t FrameId;
    uint32_t x, y, h, w;
};

struct BoundingBox {
    EIGEN_MAKE_ALIGNED_OPERATOR_NEW
    Eigen::Vector3f min = Eigen::Vector3f::Zero();
    Eigen::Vector3f max = Eigen::Vector3f::Zero();
};

struct CPUMeshData {
    std::vector<float> verts;
    std::vector<float> normals;
    std::vector<uint8_t> colors;
    std::vector<uint32_t> indices;
    bool have_reslult = false;   // (sic) spelling kept: clients read this member
    std::mutex mesh_mutex;
};

}  // namespace nerf
