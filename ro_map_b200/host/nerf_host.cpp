// nerf_host.cpp — host side of the drop-in libMON.so: nerf::NerfManagerOffline / NerfManagerOnline / NeRF /
// NeRF_Dataset with the reference's method names, argument meaning, threading and on-disk formats
// (MON/Core/src/nerf_manager.cu, nerf.cu, nerf_data.cu), implemented over the C ABI of the B200 core
// (include/mon_c.h -> libmon_b200.so).  No CUDA in this file: everything device-side is behind mon_*.
// Error behaviour follows the reference: failures print to std::cerr and exit(0) where the reference does
// (nerf_manager.cu:21-25,49-53; nerf.cu:39-42), bool returns otherwise.
#include <sys/stat.h>
#include <unistd.h>

#include <chrono>
#include <cmath>
#include <cstdlib>
#include <fstream>
#include <iostream>
#include <map>
#include <sstream>

// The viewer draws the CPU mesh with OpenGL client arrays (nerf.cu:484-507).  Where the OpenGL header exists (the RO-MAP tree:
// Pangolin / GLEW) the call is real and libMON.so needs -lGL; in a headless image it compiles to a no-op.  -DMON_NO_GL forces that.
#if !defined(MON_NO_GL) && defined(__has_include)
#if __has_include(<GL/gl.h>)
#include <GL/gl.h>
#define MON_HAVE_GL 1
#endif
#endif

#include "mon_c.h"
#include "nerf_data.h"
#include "nerf_manager.h"
#include "mesh.h"
#include "png_io.h"
#include "pose_math.h"

using std::cerr;
using std::cout;
using std::endl;
using std::string;
using std::vector;

namespace nerf {

static mon_config g_cfg;          // == NeRF_Model::ClassNetworkConfig (nerf_model.cu:16)
static bool g_cfg_loaded = false;

static void die(const string& what) {
    cerr << what << ": " << mon_last_error() << endl;
    exit(0);
}

static bool read_network_config(const string& path) {   // NeRF_Model::ReadNetworkConfig (nerf_model.cu:1272-1284)
    if (mon_config_from_json(path.c_str(), &g_cfg) != MON_OK) {
        cerr << "config file error... " << mon_last_error() << endl;
        return false;
    }
    g_cfg_loaded = true;
    return true;
}

static int detect_gpus() {
    int n = 0;
    mon_device_count(&n);
    cout << "mNumGPU: " << n << endl;
    if (n < 1) {
        cerr << "Can not Detect GPU ... " << endl;
        exit(0);
    }
    NeRF::GPUnum = n;
    cout << "Detect " << n << " GPU ..." << endl;
    return n;
}

// ------------------------------------------------------------------------------------------ NeRF_Dataset
NeRF_Dataset::~NeRF_Dataset() {
    if (mpCore) mon_dataset_destroy(mpCore);
}

// config.yaml is an OpenCV FileStorage YAML (nerf_data.cu:31-45); only flat "key: value" scalars are used
static std::map<string, double> read_flat_yaml(const string& path, bool& ok) {
    std::map<string, double> kv;
    std::ifstream f(path);
    ok = (bool)f;
    string line;
    while (std::getline(f, line)) {
        if (line.empty() || line[0] == '%' || line[0] == '#' || line.rfind("---", 0) == 0) continue;
        const size_t c = line.find(':');
        if (c == string::npos) continue;
        string key = line.substr(0, c), val = line.substr(c + 1);
        key.erase(0, key.find_first_not_of(" \t"));
        key.erase(key.find_last_not_of(" \t") + 1);
        char* end = nullptr;
        const double d = std::strtod(val.c_str(), &end);
        if (end != val.c_str()) kv[key] = d;
    }
    return kv;
}

bool NeRF_Dataset::ReadDataset(const string datasetPath) {
    bool ok = false;
    const string configPath = datasetPath + "/config.yaml";
    auto kv = read_flat_yaml(configPath, ok);
    if (!ok) {
        cerr << "Failed to open settings file at: " << configPath << endl;
        exit(0);
    }
    fx = (float)kv["Camera.fx"]; fy = (float)kv["Camera.fy"]; cx = (float)kv["Camera.cx"]; cy = (float)kv["Camera.cy"];
    H = (int)kv["Camera.H"]; W = (int)kv["Camera.W"];
    if (mbUseDepth) mfDepthScale = (float)kv["DepthMapFactor"];

    std::ifstream f(datasetPath + "/img.txt");
    string s0, s;
    std::getline(f, s0);   // skip comments
    uint32_t i = 0;
    while (std::getline(f, s)) {
        if (s.empty()) continue;
        std::stringstream ss(s);
        string stamp, name;
        ss >> stamp >> name;
        mvImagesPath.push_back(datasetPath + "/rgb/" + name);
        if (mbUseDepth) mvDepthsPath.push_back(datasetPath + "/depth/" + name);
        mvInstancesPath.push_back(datasetPath + "/instance/" + name);
        mStampToIdx[stamp] = i++;
    }
    f.close();
    f.open(datasetPath + "/groundtruth.txt");
    std::getline(f, s0);
    while (std::getline(f, s)) {
        if (s.empty()) continue;
        std::stringstream ss(s);
        string stamp;
        float tx, ty, tz, qx, qy, qz, qw;
        ss >> stamp >> tx >> ty >> tz >> qx >> qy >> qz >> qw;
        mvIamgesPose.push_back(mon_compat::pose_from_tq(tx, ty, tz, qx, qy, qz, qw));
    }
    if (mvIamgesPose.empty()) {
        cerr << "Load dataset error...No images..." << endl;
        return false;
    }
    mnImages = mvIamgesPose.size();
    return true;
}

bool NeRF_Dataset::InitDataToGPU() {   // nerf_data.cu:232-271
    if (mpCore) return true;
    if (mon_dataset_create(mGPUid, fx, fy, cx, cy, H, W, (uint32_t)mnImages, mbUseDepth ? 1 : 0, &mpCore) != MON_OK) die("dataset allocation");
    if (mbUseDepth && mbDepthRaw16 && mon_dataset_set_depth_u16(mpCore, mfDepthScale) != MON_OK) die("dataset depth format");
    return true;
}

bool NeRF_Dataset::DataToGPU() {   // nerf_data.cu:123-230
    cout << "Load Images to GPU ..." << endl;
    if (mnImages == 0) {
        cerr << "No images..." << endl;
        return false;
    }
    mbDepthRaw16 = true;
    InitDataToGPU();
    const size_t px = (size_t)H * W;
    vector<uint16_t> depth16;
    string err;
    for (size_t i = 0; i < mnImages; ++i) {
        png_io::Image rgb, inst, dep;
        if (!png_io::read(mvImagesPath[i], rgb, err) || rgb.bit_depth != 8 || rgb.width != W || rgb.height != H) {
            cerr << "Can not read image... path: " << mvImagesPath[i] << " " << err << endl;
            exit(0);
        }
        if (rgb.channels == 1) {   // cv::IMREAD_COLOR replicates gray to three channels
            vector<uint8_t> c3(px * 3);
            for (size_t p = 0; p < px; ++p) c3[3 * p] = c3[3 * p + 1] = c3[3 * p + 2] = rgb.u8[p];
            rgb.u8.swap(c3);
        }
        if (!png_io::read(mvInstancesPath[i], inst, err) || inst.bit_depth != 8 || inst.width != W || inst.height != H) {
            cerr << "Can not read image... path: " << mvInstancesPath[i] << " " << err << endl;
            exit(0);
        }
        if (inst.channels != 1) {   // instance id lives in the first channel
            vector<uint8_t> c1(px);
            for (size_t p = 0; p < px; ++p) c1[p] = inst.u8[p * inst.channels];
            inst.u8.swap(c1);
        }
        const uint16_t* dptr = nullptr;
        if (mbUseDepth) {
            if (!png_io::read(mvDepthsPath[i], dep, err) || dep.channels != 1 || dep.width != W || dep.height != H) {
                cerr << "Can not read image... path: " << mvDepthsPath[i] << " " << err << endl;
                exit(0);
            }
            // depthImg.convertTo(CV_32FC1, mfDepthScale) (nerf_data.cu:181) happens in the batch kernel, (float)u16 * mfDepthScale
            // for the pixels it picks; an 8-bit depth image is widened (same integer, same product)
            if (dep.bit_depth == 16) dptr = dep.u16.data();
            else { depth16.assign(dep.u8.begin(), dep.u8.end()); dptr = depth16.data(); }
        }
        // the PNG decoder already yields RGB (cv::imread would give BGR and the reference swaps, :163)
        if (mon_dataset_add_frame_d16(mpCore, (uint32_t)i, rgb.u8.data(), 0, inst.u8.data(), dptr, mon_compat::mat16(mvIamgesPose[i])) != MON_OK) die("frame upload");
    }
    mFrameDataNum = mnImages;
    cout << "Load Images to GPU finished, images: " << mnImages << endl;
    return true;
}

void NeRF_Dataset::FrameDataToGPU(unsigned int imgId, const string timestamp) {   // nerf_data.cu:273-339
    if (Temp_Img.empty()) { cerr << "img error ... " << endl; exit(0); }
    if (Temp_Instance.empty()) { cerr << "instance img error ... " << endl; exit(0); }
    if (mbUseDepth && Temp_Depth.empty()) { cerr << "depth img error ... " << endl; exit(0); }
    mStampToIdx[timestamp] = imgId;
    // img: 8UC3 BGR as the SLAM frontend holds it; the core stores u8 and converts in-kernel (u8 * 1/255), so the
    // caller's image is NOT converted in place as the reference does (nerf_data.cu:286-287)
    const float* dptr = mbUseDepth ? Temp_Depth.ptr<float>(0, 0) : nullptr;
    if (mon_dataset_add_frame(mpCore, imgId, Temp_Img.ptr<uint8_t>(0, 0), 1, Temp_Instance.ptr<uint8_t>(0, 0), dptr, mon_compat::mat16(Temp_Pose)) != MON_OK)
        die("frame upload");
    if (imgId >= mvIamgesPose.size()) mvIamgesPose.resize(imgId + 1, Eigen::Matrix4f::Identity());
    mvIamgesPose[imgId] = Temp_Pose;
    mFrameDataNum += 1;
}

void NeRF_Dataset::UpdateDataGPU(unsigned int CurId, unsigned int FrameNum) {   // nerf_data.cu:341-353
    if (mvTemp_Update_Pose.size() < FrameNum || CurId < FrameNum) return;
    // every object's training step holds its own mutex of this list while its batches read the poses
    // (NeRF::TrainStep, like Train_Step_Online nerf_model.cu:1676): taking all of them serialises the update
    // against every training step on this GPU
    std::vector<std::unique_lock<std::mutex>> locks;
    locks.reserve(mvUpdateMutex.size());
    for (auto& m : mvUpdateMutex) locks.emplace_back(*m);
    vector<float> flat((size_t)FrameNum * 16);
    for (unsigned int i = 0; i < FrameNum; ++i) memcpy(&flat[(size_t)i * 16], mon_compat::mat16(mvTemp_Update_Pose[i]), 64);
    const unsigned int first = CurId - FrameNum;   // `head` of the reference (nerf_data.cu:349)
    if (mon_dataset_update_poses(mpCore, first, FrameNum, flat.data()) != MON_OK) cerr << "pose update: " << mon_last_error() << endl;
    for (unsigned int i = 0; i < FrameNum && first + i < mvIamgesPose.size(); ++i) mvIamgesPose[first + i] = mvTemp_Update_Pose[i];
}

// ------------------------------------------------------------------------------------------ NeRF
int NeRF::GPUnum = -1;
int NeRF::curGPUid = -1;
int NeRF::curId = -1;

NeRF::NeRF() {   // nerf.cu:20-35: ids and round-robin GPU placement
    if (GPUnum < 1) {
        cerr << "Cannot detect GPU ..." << endl;
        exit(0);
    }
    curId += 1;
    curGPUid += 1;
    if (curGPUid >= GPUnum) curGPUid = 0;
    mId = curId;
    mGPUid = curGPUid;
    mObjTow = Eigen::Matrix4f::Zero();
}

NeRF::~NeRF() {
    if (mpCore) mon_object_destroy(mpCore);
}

bool NeRF::CreateCore() {   // NeRF_Model ctor + ResetNetwork + AllocateBatchWorkspace
    if (!g_cfg_loaded) { cerr << "network config was not read" << endl; return false; }
    if (!mpTrainData || !mpTrainData->mpCore) { cerr << "dataset is not on the GPU" << endl; return false; }
    const float bmin[3] = {mBoundingBox.min[0], mBoundingBox.min[1], mBoundingBox.min[2]};
    const float bmax[3] = {mBoundingBox.max[0], mBoundingBox.max[1], mBoundingBox.max[2]};
    if (mon_object_create(mpTrainData->mpCore, &g_cfg, 1337u, mInstanceId, mon_compat::mat16(mObjTow), bmin, bmax, &mpCore) != MON_OK) {
        cerr << "... Create Model error ... " << mon_last_error() << endl;
        return false;
    }
    return true;
}

bool NeRF::CreateModelOffline(const string path, bool useDenseDepth) {
    if (!ReadBboxOffline(path)) {
        cerr << "... Read Bbox error ..." << endl;
        exit(0);
    }
    mbUseDepth = useDenseDepth;
    if (!CreateCore()) exit(0);
    return true;
}

bool NeRF::ReadBboxOffline(const string path) {   // nerf.cu:58-118
    std::ifstream f(path);
    if (!f) {
        cerr << "Object Bbox file error..." << endl;
        return false;
    }
    string s;
    std::getline(f, s);   // skip comments
    std::getline(f, s);
    std::stringstream ss(s);
    ss >> mClass;
    mInstanceId = uint8_t(mClass);
    float num[10];   // tx,ty,tz,qx,qy,qz,qw,a1,a2,a3
    for (int i = 0; i < 10; i++) ss >> num[i];
    const Eigen::Matrix4f Two = mon_compat::pose_from_tq(num[0], num[1], num[2], num[3], num[4], num[5], num[6]);
    mObjTow = Two.inverse();
    mBoundingBox.min = Eigen::Vector3f(-num[7], -num[8], -num[9]);
    mBoundingBox.max = Eigen::Vector3f(num[7], num[8], num[9]);
    FrameIdAndBbox item;
    string stamp;
    while (std::getline(f, s)) {
        if (s.empty()) continue;
        std::stringstream ls(s);
        ls >> stamp >> item.x >> item.y >> item.h >> item.w;
        item.FrameId = mpTrainData->mStampToIdx[stamp];
        mFrameIdBbox.push_back(item);
    }
    mnBbox = mFrameIdBbox.size();
    return true;
}

void NeRF::TrainStep(int iters) {   // NeRF_Model::Train_Step / Train_Step_Online (nerf_model.cu:1630-1699)
    const auto t0 = std::chrono::steady_clock::now();
    float loss = 0.0f;
    {
        // online: pose updates of the dataset (UpdateDataGPU) wait for the step in flight and vice versa
        std::unique_lock<std::mutex> pose_lock;
        if (mpTrainData && mDataMutexIdx < mpTrainData->mvUpdateMutex.size()) pose_lock = std::unique_lock<std::mutex>(*mpTrainData->mvUpdateMutex[mDataMutexIdx]);
        if (mon_object_train(mpCore, (uint32_t)iters, &loss) != MON_OK) die("train step");
    }
    uint32_t step = 0;
    mon_object_step_count(mpCore, &step);
    mon_object_last_train_ms(mpCore, &mfLastMs);
    mfLastLoss = loss;
    mnTrainingStep = (int)step;
    const auto t1 = std::chrono::steady_clock::now();
    cout << "Id: " << mId << " train_time: " << std::chrono::duration_cast<std::chrono::milliseconds>(t1 - t0).count();
    cout << " Step: " << mnTrainingStep << " loss: " << loss << endl;
}

// GenerateMesh + TransCPUMesh (nerf_model.cu:1993-2105): 64^3 density lattice (marching_cubes.h:30), threshold 2.0 on
// the raw sigma logit, surface + normals + colours into the CPU mesh the viewer draws
void NeRF::UpdateMesh() {
    const float bmin[3] = {mBoundingBox.min[0], mBoundingBox.min[1], mBoundingBox.min[2]};
    const float bmax[3] = {mBoundingBox.max[0], mBoundingBox.max[1], mBoundingBox.max[2]};
    mesh::Extracted m;
    std::string err;
    if (!mesh::extract(mpCore, bmin, bmax, 64, 2.0f, m, err)) {
        cerr << "mesh extraction: " << err << endl;
        return;
    }
    std::unique_lock<std::mutex> lock(mCPUMeshData.mesh_mutex);
    mCPUMeshData.verts.swap(m.verts);
    mCPUMeshData.normals.swap(m.normals);
    mCPUMeshData.colors.swap(m.colors);
    mCPUMeshData.indices.swap(m.indices);
    mCPUMeshData.have_reslult = true;
}

void NeRF::SaveMesh(const std::string outname) {   // NeRF_Model::SaveMesh -> save_mesh (marching_cubes.cu:512-620)
    std::unique_lock<std::mutex> lock(mCPUMeshData.mesh_mutex);
    if (!mCPUMeshData.have_reslult) return;
    if (!mesh::save_ply(outname, mCPUMeshData.verts, mCPUMeshData.normals, mCPUMeshData.colors, mCPUMeshData.indices))
        cerr << "Failed to open " << outname << " for writing." << endl;
}

void NeRF::TrainOffline(const int iterations) {   // nerf.cu:120-153
    const auto start = std::chrono::steady_clock::now();
    if (mon_object_set_bboxes(mpCore, reinterpret_cast<const mon_bbox2d*>(mFrameIdBbox.data()), (uint32_t)mFrameIdBbox.size()) != MON_OK) die("bbox upload");
    mnCoreBbox = mFrameIdBbox.size();
    const auto allocate_time = std::chrono::steady_clock::now();
    cout << "allocate_time: " << std::chrono::duration_cast<std::chrono::milliseconds>(allocate_time - start).count() << std::endl;
    for (int i = 1; i <= iterations; i++) {
        TrainStep(500);
        if (i % 2 == 0) UpdateMesh();
    }
    mkdir("./output", 0755);
    SaveMesh("./output/" + std::to_string(mId) + ".ply");
    cout << "Training completed, press Ctrl+C to exit" << endl;
}

void NeRF::SetAttributes(const int Class, const Eigen::Matrix4f& ObjTow, const BoundingBox& BBox, size_t maxnumBbox) {   // nerf.cu:155-177
    mClass = Class;
    mInstanceId = uint8_t(Class);
    mObjTow = ObjTow;
    mBoundingBox = BBox;
    const float k = (Class == 41 || Class == 73) ? 1.2f : 1.1f;   // appropriately expand the 3D bounding box
    mBoundingBox.max = k * mBoundingBox.max;
    mBoundingBox.min = k * mBoundingBox.min;
    mFrameIdBbox.resize(maxnumBbox);
    mnBbox = 0;
}

bool NeRF::CreateModelOnline(bool useSparseDepth, int Iterations) {
    mbUseDepth = useSparseDepth;
    mnIteration = Iterations;
    return true;
}

void NeRF::TrainOnline() {   // nerf.cu:187-253
    if (!CreateCore()) exit(0);
    // instantiate the iteration graphs now, while the object waits for its first boxes: done at the first training step
    // (all objects of a scene reach it together) it held the driver for ~100 ms and the frontend thread's keyframe upload with it
    if (mon_object_prepare_train(mpCore, (uint32_t)mnIteration) != MON_OK) cerr << "graph preparation: " << mon_last_error() << endl;
    int train_step_count = 0;
    while (1) {
        int train_step = 0;
        {
            std::unique_lock<std::mutex> lock(mUpdateBbox);
            if (mnBbox == mnCoreBbox && !mbFinishRequested) mCond.wait(lock);   // no update, wait
            if (mnBbox > mnCoreBbox) {
                const mon_bbox2d* fresh = reinterpret_cast<const mon_bbox2d*>(mFrameIdBbox.data() + mnCoreBbox);
                const uint32_t n_fresh = (uint32_t)(mnBbox - mnCoreBbox);
                if (mon_object_add_bboxes(mpCore, fresh, n_fresh) != MON_OK) {
                    // a box outside the image or on a keyframe that has not been ingested: drop that box and keep the
                    // SLAM process alive (a training thread must not exit() the whole system over one detection)
                    for (uint32_t k = 0; k < n_fresh; ++k)
                        if (mon_object_add_bboxes(mpCore, fresh + k, 1) != MON_OK) cerr << "Id: " << mId << " box dropped: " << mon_last_error() << endl;
                }
                mnCoreBbox = mnBbox;
                train_step = mnTrainStep;
                mnTrainStep = 0;
            }
        }
        if (mnCoreBbox > 10) {
            for (int i = 0; i < train_step; i++) {
                TrainStep(mnIteration);
                train_step_count += 1;
                if (train_step_count % 2 == 0) UpdateMesh();
            }
        }
        if (CheckFinish()) break;
        usleep(3000);
    }
    if (mnCoreBbox > 0) TrainStep(mnIteration);   // last time
    UpdateMesh();
    cout << "Id: " << mId << " finished! " << endl;
}

void NeRF::UpdateFrameBBox(const vector<nerf::FrameIdAndBbox>& vFrameBbox, const int train_step) {   // nerf.cu:406-422
    std::unique_lock<std::mutex> lock(mUpdateBbox);
    if (mnBbox + vFrameBbox.size() > mFrameIdBbox.size()) mFrameIdBbox.resize(mnBbox + vFrameBbox.size());
    for (size_t i = 0; i < vFrameBbox.size(); i++) mFrameIdBbox[mnBbox + i] = vFrameBbox[i];
    mnBbox += vFrameBbox.size();
    mnTrainStep = train_step;
    mCond.notify_all();
}

bool NeRF::CheckFinish() {
    std::unique_lock<std::mutex> lock(mFinishMutex);
    return mbFinishRequested;
}

void NeRF::RequestFinish() {
    {
        std::unique_lock<std::mutex> lock(mFinishMutex);
        mbFinishRequested = true;
    }
    std::unique_lock<std::mutex> lock(mUpdateBbox);
    mCond.notify_all();
}

// float image -> PNG sample exactly like cv::Mat::convertTo(CV_8U / CV_16U, scale): round half to even, saturate
static inline uint8_t to_u8(float v, float scale) {
    const float r = std::nearbyint(v * scale);
    return (uint8_t)(r < 0.0f ? 0.0f : r > 255.0f ? 255.0f : r);
}
static inline uint16_t to_u16(float v, float scale) {
    const float r = std::nearbyint(v * scale);
    return (uint16_t)(r < 0.0f ? 0.0f : r > 65535.0f ? 65535.0f : r);
}
static void write_view_pngs(const string& img_path, const string& depth_path, const string& mask_path, int w, int h,
                            const vector<float>& rgb, const vector<float>& depth, const vector<float>* mask) {
    const size_t n = (size_t)w * h;
    vector<uint8_t> rgb8(n * 3);
    vector<uint16_t> d16(n);
    for (size_t p = 0; p < n * 3; ++p) rgb8[p] = to_u8(rgb[p], 255.0f);
    for (size_t p = 0; p < n; ++p) d16[p] = to_u16(depth[p], 20000.0f);   // "*20000, looks obvious" (nerf.cu:346)
    png_io::write(img_path, w, h, 3, 8, rgb8.data());
    png_io::write(depth_path, w, h, 1, 16, d16.data());
    if (mask) {
        vector<uint8_t> mask8(n);
        for (size_t p = 0; p < n; ++p) mask8[p] = to_u8((*mask)[p], 255.0f);
        png_io::write(mask_path, w, h, 1, 8, mask8.data());
    }
}
// "stamp x y h w tx ty tz qx qy qz qw" with the object-centric camera pose Toc = ObjTow * Twc (nerf.cu:331-336)
static void write_pose_line(std::ofstream& f, const string& stamp, const FrameIdAndBbox& b, const Eigen::Matrix4f& ObjTow, const Eigen::Matrix4f& Twc) {
    float Toc[16], q[4];
    pose_math::mul44(mon_compat::mat16(ObjTow), mon_compat::mat16(Twc), Toc);
    pose_math::rot_to_quat(Toc, q);
    f << stamp << " " << b.x << " " << b.y << " " << b.h << " " << b.w << " " << Toc[12] << " " << Toc[13] << " " << Toc[14] << " "
      << q[0] << " " << q[1] << " " << q[2] << " " << q[3] << endl;
}

// NeRF::RenderTestImg (nerf.cu:255-404): <out>/<id>/{test_img,test_depth,test_mask}/<stamp>.png + test.txt, train.txt,
// the 60-view 360-degree video {video_img,video_depth}/<i>.png (NeRF_Model::RenderVideo, nerf_model.cu:1832-1991) and
// obj.ply.  PNG scaling as the reference: RGB x255 8U, z-depth x20000 16U, mask x255 8U.
void NeRF::RenderTestImg(const string out_path, const vector<string>& timestamp, const vector<Eigen::Matrix4f>& testTwc,
                         const vector<FrameIdAndBbox>& testBbox, const float radius) {
    const string folder = out_path + "/" + std::to_string(mId);
    mkdir(out_path.c_str(), 0755);
    for (const char* sub : {"", "/test_img", "/test_depth", "/test_mask", "/video_img", "/video_depth"}) mkdir((folder + sub).c_str(), 0755);

    std::ofstream f(folder + "/test.txt");
    f << std::fixed;
    f << "#stamp  box.x  box.y  box.h  box.w  tx  ty  tz  qx  qy  qz  qw (object-centric)" << endl;
    cout << "Render Object " << mId << " test imgs to " << folder << "/test_img ... please wait..." << endl;
    for (size_t i = 0; i < timestamp.size() && i < testTwc.size() && i < testBbox.size(); ++i) {
        const FrameIdAndBbox& b = testBbox[i];
        const string& stamp = timestamp[i];
        write_pose_line(f, stamp, b, mObjTow, testTwc[i]);
        if (b.h == 0 || b.w == 0) continue;
        const size_t n = (size_t)b.h * b.w;
        vector<float> rgb(n * 3), depth(n), mask(n);
        mon_bbox2d box = {b.FrameId, b.x, b.y, b.h, b.w};
        if (mon_object_render(mpCore, box, mon_compat::mat16(testTwc[i]), 1, nullptr, rgb.data(), depth.data(), mask.data()) != MON_OK) {
            cerr << "render: " << mon_last_error() << endl;
            continue;
        }
        write_view_pngs(folder + "/test_img/" + stamp + ".png", folder + "/test_depth/" + stamp + ".png", folder + "/test_mask/" + stamp + ".png",
                        (int)b.w, (int)b.h, rgb, depth, &mask);
    }
    f.close();

    // training views (nerf.cu:364-396): class + box half-extent, then one pose line per 2-D box
    f.open(folder + "/train.txt");
    f << std::fixed;
    f << "#class Bbox" << endl;
    f << mClass << " ";
    for (int i = 0; i < 3; ++i) f << mBoundingBox.max[i] << " ";
    f << endl;
    f << "#stamp box.x box.y box.h box.w  tx  ty  tz  qx  qy  qz  qw (object-centric)" << endl;
    if (mpTrainData) {
        std::map<uint32_t, string> idx_to_stamp;
        for (const auto& kv : mpTrainData->mStampToIdx) idx_to_stamp.emplace(kv.second, kv.first);
        for (size_t i = 0; i < mnBbox && i < mFrameIdBbox.size(); ++i) {
            const FrameIdAndBbox& b = mFrameIdBbox[i];
            auto it = idx_to_stamp.find(b.FrameId);
            if (it == idx_to_stamp.end() || b.FrameId >= mpTrainData->mvIamgesPose.size()) continue;
            write_pose_line(f, it->second, b, mObjTow, mpTrainData->mvIamgesPose[b.FrameId]);
        }
    }
    f.close();

    RenderVideo(folder + "/video_img", folder + "/video_depth", radius);

    cout << "Save Object Mesh ... please wait..." << endl;
    UpdateMesh();
    SaveMesh(folder + "/obj.ply");
}

// NeRF_Model::RenderVideo (nerf_model.cu:1832-1991): 60 turn-table views at elevation 30 degrees, the central half of
// the image (x = W/4, y = H/4, w = W/2, h = H/2), camera poses given directly in the object frame.
void NeRF::RenderVideo(const string img_path_folder, const string depth_path_folder, const float radius) {
    if (!mpCore || !mpTrainData) return;
    cout << "Render Object " << mId << " 360 video imgs to " << img_path_folder << " ... please wait..." << endl;
    const int theta_num = 60;
    const float theta = 360 / float(theta_num), phi = 30;
    float cur_theta = 0.0f;
    mon_bbox2d box = {0u, (uint32_t)(mpTrainData->W / 4), (uint32_t)(mpTrainData->H / 4), (uint32_t)(mpTrainData->H / 2), (uint32_t)(mpTrainData->W / 2)};
    if (box.h == 0 || box.w == 0) return;
    const size_t n = (size_t)box.h * box.w;
    vector<float> rgb(n * 3), depth(n), mask(n);
    for (int i = 0; i < theta_num; ++i) {
        cur_theta += theta;
        float Toc[16];
        pose_math::turntable_toc(cur_theta, phi, radius, Toc);
        if (mon_object_render_object_centric(mpCore, box, Toc, 1, nullptr, rgb.data(), depth.data(), mask.data()) != MON_OK) {
            cerr << "render video: " << mon_last_error() << endl;
            return;
        }
        write_view_pngs(img_path_folder + "/" + std::to_string(i) + ".png", depth_path_folder + "/" + std::to_string(i) + ".png", "",
                        (int)box.w, (int)box.h, rgb, depth, nullptr);
    }
}

vector<Eigen::Matrix4f> NeRF::GetTwc() {
    vector<Eigen::Matrix4f> Twc;
    for (size_t i = 0; i < mnBbox && i < mFrameIdBbox.size(); i++) Twc.push_back(mpTrainData->mvIamgesPose[mFrameIdBbox[i].FrameId]);
    return Twc;
}
BoundingBox NeRF::GetBoundingBox() { return mBoundingBox; }
Eigen::Matrix4f NeRF::GetObjTow() { return mObjTow; }
CPUMeshData& NeRF::GetCPUMeshData() { return mCPUMeshData; }
vector<FrameIdAndBbox> NeRF::GetFrameIdAndBBox() { return mFrameIdBbox; }
// Called by the viewer thread every frame (src/MapDrawer.cc:396, MON/main.cpp): draws the latest CPU mesh from client arrays —
// positions, 1-ring normals, u8 colours, the reference's triangle winding — and never waits for a mesh update in flight
// (try-lock, the frame is skipped instead; nerf.cu:488-490).
void NeRF::DrawCPUMesh() {
#ifdef MON_HAVE_GL
    std::unique_lock<std::mutex> lock(mCPUMeshData.mesh_mutex, std::try_to_lock);
    if (!lock.owns_lock() || !mCPUMeshData.have_reslult) return;
    const CPUMeshData& m = mCPUMeshData;
    static const GLenum kArrays[3] = {GL_VERTEX_ARRAY, GL_NORMAL_ARRAY, GL_COLOR_ARRAY};
    for (GLenum a : kArrays) glEnableClientState(a);
    glVertexPointer(3, GL_FLOAT, 0, m.verts.data());
    glNormalPointer(GL_FLOAT, 0, m.normals.data());
    glColorPointer(3, GL_UNSIGNED_BYTE, 0, m.colors.data());
    glDrawElements(GL_TRIANGLES, (GLsizei)m.indices.size(), GL_UNSIGNED_INT, m.indices.data());
    for (GLenum a : kArrays) glDisableClientState(a);
#endif
}
void NeRF::DrawMesh() { DrawCPUMesh(); }   // the reference's VBO variant is dead code (nerf.cu:140-145 uses the CPU mesh only)

// ------------------------------------------------------------------------------------------ NerfManagerOffline
NerfManagerOffline::NerfManagerOffline(const string datasetPath, const string networkConfigFile, bool useDenseDepth)
    : msNetworkConfigFile(networkConfigFile), msDatasetPath(datasetPath), mbUseDenseDepth(useDenseDepth) {}

bool NerfManagerOffline::Init() {
    mNumGPU = detect_gpus();
    if (!read_network_config(msNetworkConfigFile)) {
        cerr << "Read Network Config error..." << endl;
        exit(0);
    }
    return true;
}

bool NerfManagerOffline::ReadDataset() {   // one dataset replica per GPU (nerf_manager.cu:40-62)
    std::vector<std::thread> threads_data;
    for (int i = 0; i < mNumGPU; i++) {
        auto pDataset = std::make_shared<NeRF_Dataset>(mbUseDenseDepth);
        mvpDataset.push_back(pDataset);
        pDataset->mGPUid = i;
        if (!pDataset->ReadDataset(msDatasetPath)) {
            cerr << "Read Train Data error..." << endl;
            exit(0);
        }
        if (i == 0) threads_data.emplace_back(std::thread(&NeRF_Dataset::DataToGPU, pDataset));
    }
    for (auto& t : threads_data) t.join();
    // the other replicas are copied GPU-to-GPU over NVLink instead of being decoded and uploaded once per GPU
    for (int i = 1; i < mNumGPU; i++) {
        mvpDataset[i]->mbDepthRaw16 = mvpDataset[0]->mbDepthRaw16;
        mvpDataset[i]->InitDataToGPU();
        if (mon_dataset_clone_from_peer(mvpDataset[i]->mpCore, mvpDataset[0]->mpCore) != MON_OK) die("dataset replication");
        mvpDataset[i]->mFrameDataNum = mvpDataset[0]->mFrameDataNum;
    }
    return true;
}

bool NerfManagerOffline::CreateNeRF(const string objectFile) {
    std::ifstream file(objectFile);
    if (!file) {
        cerr << "object file error..." << endl;
        return false;
    }
    auto NeRFInstance = std::make_shared<NeRF>();
    mvpNeRFs.push_back(NeRFInstance);
    NeRFInstance->mpTrainData = mvpDataset[NeRFInstance->mGPUid];
    if (!NeRFInstance->CreateModelOffline(objectFile, mbUseDenseDepth)) {
        cerr << "Create NeRF error ..." << endl;
        exit(0);
    }
    mvThreads.emplace_back(std::thread(&NeRF::TrainOffline, NeRFInstance, mnTrainSteps));
    return true;
}

bool NerfManagerOffline::WaitThreadsEnd() {
    if (mvThreads.empty()) return false;
    for (std::thread& t : mvThreads) t.join();
    mvThreads.clear();
    return true;
}

std::shared_ptr<NeRF> NerfManagerOffline::GetNeRF(int idx) {
    if (idx < 0 || (size_t)idx >= mvpNeRFs.size()) {
        cerr << "NeRF Idx error ... " << endl;
        exit(0);
    }
    return mvpNeRFs[idx];
}
vector<std::shared_ptr<NeRF>> NerfManagerOffline::GetAllNeRF() { return mvpNeRFs; }
vector<Eigen::Matrix4f> NerfManagerOffline::GetAllTwc() { return mvpDataset[0]->mvIamgesPose; }
void NerfManagerOffline::GetIntrinsics(float& fx, float& fy, float& cx, float& cy) {
    fx = mvpDataset[0]->fx; fy = mvpDataset[0]->fy; cx = mvpDataset[0]->cx; cy = mvpDataset[0]->cy;
}

// ------------------------------------------------------------------------------------------ NerfManagerOnline
NerfManagerOnline::NerfManagerOnline(const string network_config_file, bool UseSparseDepth, int TrainStepIterations)
    : mNetworkConfigFile(network_config_file), mbUseSparseDepth(UseSparseDepth), mnTrainStepIterations(TrainStepIterations) {}

bool NerfManagerOnline::Init() {
    mNumGPU = detect_gpus();
    if (!read_network_config(mNetworkConfigFile)) {
        cerr << "Read Network Config error..." << endl;
        exit(0);
    }
    return true;
}

void NerfManagerOnline::DatasetInit(float fx, float fy, float cx, float cy, int H, int W, size_t imgs) {
    for (int i = 0; i < mNumGPU; i++) {
        auto pDataset = std::make_shared<NeRF_Dataset>(mbUseSparseDepth);
        mvpDataset.push_back(pDataset);
        pDataset->mGPUid = i;
        pDataset->fx = fx; pDataset->fy = fy; pDataset->cx = cx; pDataset->cy = cy;
        pDataset->H = H; pDataset->W = W;
        pDataset->mfDepthScale = 1.0f;
        pDataset->mnImages = imgs;
        pDataset->InitDataToGPU();
    }
}

void NerfManagerOnline::NewFrameToDataset(unsigned int imgId, const string timestamp, cv::Mat& img, cv::Mat& instance, const cv::Mat& depth_img,
                                          const Eigen::Matrix4f& pose) {
    // the reference uploads the frame once per GPU, from one host thread each (nerf_manager.cu:189-217).  Here it crosses
    // PCIe ONCE, to GPU 0 (asynchronous DMA out of pinned staging memory), and every other replica is filled device to
    // device over NVLink on that dataset's own upload stream, ordered behind the upload by an event: nothing below waits
    // for a copy, so the SLAM thread pays one staging memcpy whatever the number of GPUs
    {
        auto& d = mvpDataset[0];
        d->Temp_Img = img; d->Temp_Instance = instance; d->Temp_Depth = depth_img; d->Temp_Pose = pose;
        d->FrameDataToGPU(imgId, timestamp);
    }
    for (int i = 1; i < mNumGPU; i++) {
        auto& d = mvpDataset[i];
        if (mon_dataset_copy_frame_from_peer(d->mpCore, mvpDataset[0]->mpCore, imgId) != MON_OK) die("frame replication");
        d->mStampToIdx[timestamp] = imgId;
        if (imgId >= d->mvIamgesPose.size()) d->mvIamgesPose.resize(imgId + 1, Eigen::Matrix4f::Identity());
        d->mvIamgesPose[imgId] = pose;
        d->mFrameDataNum += 1;
    }
}

void NerfManagerOnline::UpdateDataset(unsigned int CurId, unsigned int FrameNum, const vector<Eigen::Matrix4f>& Poses) {
    for (int i = 0; i < mNumGPU; i++) {
        mvpDataset[i]->mvTemp_Update_Pose = Poses;
        mvpDataset[i]->UpdateDataGPU(CurId, FrameNum);
    }
}

size_t NerfManagerOnline::CreateNeRF(const int Class, const Eigen::Matrix4f& ObjTow, const nerf::BoundingBox& BoundingBox) {
    auto NeRFInstance = std::make_shared<NeRF>();
    const size_t idx = mvpNeRFs.size();
    mvpNeRFs.push_back(NeRFInstance);
    NeRFInstance->mpTrainData = mvpDataset[NeRFInstance->mGPUid];
    NeRFInstance->mDataMutexIdx = NeRFInstance->mpTrainData->mvUpdateMutex.size();
    NeRFInstance->mpTrainData->mvUpdateMutex.emplace_back(new std::mutex());
    NeRFInstance->SetAttributes(Class, ObjTow, BoundingBox, mvpDataset[0]->mnImages);
    if (!NeRFInstance->CreateModelOnline(mbUseSparseDepth, mnTrainStepIterations)) {
        cerr << "Create NeRF error ..." << endl;
        exit(0);
    }
    mvThreads.emplace_back(std::thread(&NeRF::TrainOnline, NeRFInstance));
    return idx;
}

bool NerfManagerOnline::WaitThreadsEnd() {
    if (mvThreads.empty()) return false;
    for (auto& p : mvpNeRFs) p->RequestFinish();
    for (std::thread& t : mvThreads) t.join();
    mvThreads.clear();
    cout << "All NeRF threads completed ..." << endl;
    return true;
}

void NerfManagerOnline::RenderNeRFsTest(const string out_path, const size_t Idx, const vector<string>& timestamp, const vector<FrameIdAndBbox>& vBbox,
                                        const vector<Eigen::Matrix4f>& vTwc, const float radius) {
    if (mvpNeRFs.empty()) return;
    mvpNeRFs[Idx]->RenderTestImg(out_path, timestamp, vTwc, vBbox, radius);
}

int NerfManagerOnline::GetFrameIdx(double timastamp) {
    const string stamp = std::to_string(timastamp);
    auto& m = mvpDataset[0]->mStampToIdx;
    auto it = m.find(stamp);
    return it == m.end() ? -1 : int(it->second);
}

void NerfManagerOnline::UpdateNeRFBbox(const size_t idx, const vector<nerf::FrameIdAndBbox>& vFrameBbox, const int train_step) {
    if (vFrameBbox.empty()) return;
    mvpNeRFs[idx]->UpdateFrameBBox(vFrameBbox, train_step);
}

void NerfManagerOnline::DrawMesh(size_t idx) {
    if (mvpNeRFs.empty() || idx > (mvpNeRFs.size() - 1)) return;
    mvpNeRFs[idx]->DrawCPUMesh();
}

}  // namespace nerf
