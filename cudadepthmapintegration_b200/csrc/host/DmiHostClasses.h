// VTK-free C++ mirrors of the reference's two operators, over the C ABI of include/dmi_b200.h:
//   dmihost::CudaReconstructionFilter  ~  vtkCudaReconstructionFilter (Reconstruction/vtkCudaReconstructionFilter.h:48-120)
//   dmihost::MeshColoration            ~  MeshColoration              (Coloration/MeshColoration.h:42-62)
// Same setter names, same argument meaning, same error behaviour (messages on std::cerr, 0 / false on
// the reference's error paths).  vtkImageData / vtkPolyData become plain arrays; the list files, .krtd
// and .vti files are read by DmiHelper.h / DmiVti.h.  Header-only; link libdmi_b200.so.
#pragma once
#include <algorithm>
#include <chrono>
#include <cstdint>
#include <iostream>
#include <string>
#include <vector>

#include "../../../include/dmi_b200.h"
#include "DmiHelper.h"
#include "DmiVti.h"

namespace dmihost {

class CudaReconstructionFilter
{
public:
  explicit CudaReconstructionFilter(int device = 0) { if (dmi_create(device, &ctx_) != DMI_OK) std::cerr << dmi_last_error(nullptr) << std::endl; }
  ~CudaReconstructionFilter() { dmi_destroy(ctx_); dmi_group_destroy(group_); }
  CudaReconstructionFilter(const CudaReconstructionFilter&) = delete;
  void operator=(const CudaReconstructionFilter&) = delete;

  // vtkSetMacro members of vtkCudaReconstructionFilter.h:57-77
  void SetRayPotentialThickness(double v) { RayPotentialThickness = v; }
  void SetRayPotentialRho(double v) { RayPotentialRho = v; }
  void SetRayPotentialEta(double v) { RayPotentialEta = v; }
  void SetRayPotentialDelta(double v) { RayPotentialDelta = v; }
  void SetThresholdBestCost(double v) { ThresholdBestCost = v; }
  void SetFilePathKRTD(const std::string& p) { FilePathKRTD = p; }
  void SetFilePathVTI(const std::string& p) { FilePathVTI = p; }
  void SetGridMatrix(const double m[16]) { std::copy(m, m + 16, GridMatrix); }
  double GetExecutionTime() const { return ExecutionTime; }
  // new (the reference is single-GPU): run on several GPUs of the box; the z-layers of the grid are dealt to them and the
  // views exchanged over NVLink inside libdmi_b200 (dmi_group_*); the output is bit-identical to one GPU's
  bool SetDevices(const std::vector<int>& devices)
  {
    dmi_group_destroy(group_);
    group_ = nullptr;
    if (devices.size() < 2) return true;
    if (dmi_group_create(devices.data(), (int)devices.size(), &group_) != DMI_OK)
    { std::cerr << dmi_last_error(nullptr) << std::endl; group_ = nullptr; return false; }
    return true;
  }

  // stands for SetInputData(vtkImageData*): GetDimensions / GetOrigin / GetSpacing (.cxx:121-126); POINT dims
  void SetInputGrid(const int dims[3], const double origin[3], const double spacing[3])
  {
    for (int a = 0; a < 3; a++) { Dims[a] = dims[a]; Origin[a] = origin[a]; Spacing[a] = spacing[a]; }
    hasGrid_ = true;
  }

  // RequestData (.cxx:96-151): 1 on success, 0 on the reference's error paths.  The reference ignores Compute()'s
  // result (.cxx:144) and would hand on a zero volume after a failed read; here a failed Compute() (unreadable or
  // corrupt file, size mismatch, CUDA / dmi error) makes Update() return 0 and leaves no output to write.
  int Update()
  {
    ExecutionTime = -1;
    const auto start = std::chrono::steady_clock::now();
    if (!ctx_) return 0;
    if (FilePathKRTD.empty() || FilePathVTI.empty() || !hasGrid_)
    {
      std::cerr << "Error, some inputs have not been set." << std::endl;
      return 0;
    }
    const size_t cells = (size_t)(Dims[0] - 1) * (Dims[1] - 1) * (Dims[2] - 1);
    Output.assign(cells, 0.0);                                      // FillComponent(0, 0), .cxx:133
    if (RayPotentialRho == 0 && RayPotentialThickness == 0)
    {
      std::cerr << "Error : Ray potential Rho or Thickness or both have not been set" << std::endl;
      return 0;
    }
    if (Compute() != 0) { Output.clear(); return 0; }
    ExecutionTime = std::chrono::duration<double>(std::chrono::steady_clock::now() - start).count();
    return 1;
  }

  // the "reconstruction_scalar" cell array, VTK cell order
  const std::vector<double>& GetOutput() const { return Output; }

private:
  // Compute (.cxx:155-179) + the loop of ProcessDepthMap<double> (CudaReconstruction.cu:343-365)
  int Compute()
  {
    const std::vector<std::string> vtiList = help::ExtractAllFilePath(FilePathVTI.c_str());
    const std::vector<std::string> krtdList = help::ExtractAllFilePath(FilePathKRTD.c_str());
    if (vtiList.size() == 0 || krtdList.size() < vtiList.size())
    {
      std::cerr << "Error : There is no enough vti files, please check your vtiList.txt and krtdList.txt" << std::endl;
      return -1;
    }
    DepthMapImage img;
    std::string err;
    if (!ReadVti(vtiList[0], img, err)) { std::cerr << err << std::endl; return -1; }
    const int dd[2] = {img.W, img.H};
    if (group_) return ComputeOnGroup(vtiList, krtdList, dd);
    if (dmi_initialize(ctx_, GridMatrix, Dims, Origin, Spacing, RayPotentialThickness, RayPotentialRho,
                       RayPotentialEta, RayPotentialDelta, dd) != DMI_OK ||
        dmi_volume_begin(ctx_, nullptr, DMI_F64) != DMI_OK)      // Output was just zero-filled: nothing to upload
    { std::cerr << dmi_last_error(ctx_) << std::endl; return -1; }
    const size_t npix = (size_t)img.W * img.H;
    const size_t batch = std::max<size_t>(1, std::min<size_t>(32, (256u << 20) / (npix * 8)));
    std::vector<double> depth, cost, K, RT;
    std::cout << "START CUDA ON " << vtiList.size() << " Depth map" << std::endl;
    for (size_t v0 = 0; v0 < vtiList.size(); v0 += batch)
    {
      const size_t n = std::min(batch, vtiList.size() - v0);
      depth.clear(); cost.clear(); K.assign(16 * n, 0.0); RT.assign(16 * n, 0.0);
      bool haveCost = true;
      for (size_t v = 0; v < n; v++)
      {
        if (!ReadVti(vtiList[v0 + v], img, err)) { std::cerr << err << std::endl; return -1; }
        if ((size_t)img.W * img.H != npix) { std::cerr << vtiList[v0 + v] << ": depth map size differs from the first one" << std::endl; return -1; }
        if (!help::ReadKrtdFile(krtdList[v0 + v], &K[16 * v], &RT[16 * v])) return -1;
        depth.insert(depth.end(), img.depths.begin(), img.depths.end());
        // ApplyDepthThresholdFilter is a no-op when the array sizes differ (ReconstructionData.cxx:156-157)
        if (img.bestCost.size() == npix) cost.insert(cost.end(), img.bestCost.begin(), img.bestCost.end());
        else haveCost = false;
      }
      if (!haveCost && !cost.empty())
      {
        std::cerr << "Error : 'Best Cost Values' present in some depth maps of a batch only" << std::endl;
        return -1;
      }
      if (dmi_volume_integrate_host(ctx_, (int)n, depth.data(), haveCost ? cost.data() : nullptr, ThresholdBestCost,
                                    K.data(), RT.data()) != DMI_OK)
      { std::cerr << dmi_last_error(ctx_) << std::endl; return -1; }
    }
    if (dmi_volume_end(ctx_, Output.data()) != DMI_OK) { std::cerr << dmi_last_error(ctx_) << std::endl; return -1; }
    return 0;
  }

  // several GPUs: every view is read first (each GPU then uploads only the views it owns), one call integrates them
  int ComputeOnGroup(const std::vector<std::string>& vtiList, const std::vector<std::string>& krtdList, const int dd[2])
  {
    if (dmi_group_initialize(group_, GridMatrix, Dims, Origin, Spacing, RayPotentialThickness, RayPotentialRho, RayPotentialEta,
                             RayPotentialDelta, dd) != DMI_OK)
    { std::cerr << dmi_group_last_error(group_) << std::endl; return -1; }
    const size_t npix = (size_t)dd[0] * dd[1], n = vtiList.size();
    std::vector<double> depth, cost, K(16 * n, 0.0), RT(16 * n, 0.0);
    depth.reserve(n * npix);
    bool haveCost = true;
    std::cout << "START CUDA ON " << n << " Depth map" << std::endl;
    for (size_t v = 0; v < n; v++)
    {
      DepthMapImage img;
      std::string err;
      if (!ReadVti(vtiList[v], img, err)) { std::cerr << err << std::endl; return -1; }
      if ((size_t)img.W * img.H != npix) { std::cerr << vtiList[v] << ": depth map size differs from the first one" << std::endl; return -1; }
      if (!help::ReadKrtdFile(krtdList[v], &K[16 * v], &RT[16 * v])) return -1;
      depth.insert(depth.end(), img.depths.begin(), img.depths.end());
      if (img.bestCost.size() == npix) cost.insert(cost.end(), img.bestCost.begin(), img.bestCost.end());
      else haveCost = false;
    }
    if (!haveCost && !cost.empty())
    { std::cerr << "Error : 'Best Cost Values' present in some depth maps only" << std::endl; return -1; }
    if (dmi_group_process_depth_maps(group_, (int)n, depth.data(), haveCost ? cost.data() : nullptr, ThresholdBestCost, K.data(),
                                     RT.data(), Output.data(), DMI_F64) != DMI_OK)
    { std::cerr << dmi_group_last_error(group_) << std::endl; return -1; }
    return 0;
  }

  dmi_ctx* ctx_ = nullptr;
  dmi_group* group_ = nullptr;
  bool hasGrid_ = false;
  int Dims[3] = {0, 0, 0};
  double Origin[3] = {0, 0, 0}, Spacing[3] = {1, 1, 1};
  double GridMatrix[16] = {1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1};
  double RayPotentialRho = 0, RayPotentialThickness = 0, RayPotentialEta = 0, RayPotentialDelta = 0;
  double ThresholdBestCost = 0, ExecutionTime = -1;
  std::string FilePathKRTD, FilePathVTI;
  std::vector<double> Output;
};

class MeshColoration
{
public:
  // MeshColoration(vtkPolyData* mesh, std::string vti, std::string krtd) (MeshColoration.cxx:52-72):
  // deep copy of the mesh's points (float32 xyz, vtkPoints' default storage); every view is loaded now.
  MeshColoration(const std::vector<float>& xyz, const std::string& vti, const std::string& krtd, int device = 0)
  {
    if (dmi_create(device, &ctx_) != DMI_OK) std::cerr << dmi_last_error(nullptr) << std::endl;
    points_ = xyz;
    hasInput_ = true;
    const std::vector<std::string> vtiList = help::ExtractAllFilePath(vti.c_str());
    const std::vector<std::string> krtdList = help::ExtractAllFilePath(krtd.c_str());
    if (krtdList.size() < vtiList.size())
    {
      std::cerr << "Error, not enough krtd file for each vti file" << std::endl;
      return;
    }
    for (size_t id = 0; id < vtiList.size(); id++)
    {
      DepthMapImage img;
      std::string err;
      if (!ReadVti(vtiList[id], img, err)) { std::cerr << err << std::endl; views_ = 0; return; }
      if (id == 0) { W_ = img.W; H_ = img.H; }
      if (img.color.size() != (size_t)W_ * H_ * 3)
      { std::cerr << "Error, no 'Color' array exists" << std::endl; views_ = 0; return; }     // ReconstructionData.cxx:97-101
      colors_.insert(colors_.end(), img.color.begin(), img.color.end());
      K_.resize(16 * (id + 1)); RT_.resize(16 * (id + 1));
      if (!help::ReadKrtdFile(krtdList[id], &K_[16 * id], &RT_[16 * id])) { views_ = 0; return; }
      views_ = (int)id + 1;
    }
  }
  ~MeshColoration() { dmi_destroy(ctx_); dmi_group_destroy(group_); }
  MeshColoration(const MeshColoration&) = delete;
  void operator=(const MeshColoration&) = delete;

  void SetInput(const std::vector<float>& xyz) { points_ = xyz; hasInput_ = true; }
  // new: colour on several GPUs (points sharded by index, colour images exchanged over NVLink; dmi_group_colorize)
  bool SetDevices(const std::vector<int>& devices)
  {
    dmi_group_destroy(group_);
    group_ = nullptr;
    if (devices.size() < 2) return true;
    if (dmi_group_create(devices.data(), (int)devices.size(), &group_) != DMI_OK)
    { std::cerr << dmi_last_error(nullptr) << std::endl; group_ = nullptr; return false; }
    return true;
  }

  // ProcessColoration (MeshColoration.cxx:98-199)
  bool ProcessColoration()
  {
    if (!ctx_ || !hasInput_ || views_ == 0)
    {
      std::cerr << "Error when input has been set or during reading vti/krtd file path" << std::endl;
      return false;
    }
    const size_t P = points_.size() / 3;
    MeanColoration.assign(3 * P, 0); MedianColoration.assign(3 * P, 0); NbProjectedDepthMap.assign(P, 0);
    if (group_)
    {
      if (dmi_group_colorize(group_, P, points_.data(), DMI_F32, views_, colors_.data(), K_.data(), RT_.data(), W_, H_,
                             MeanColoration.data(), MedianColoration.data(), NbProjectedDepthMap.data()) != DMI_OK)
      { std::cerr << dmi_group_last_error(group_) << std::endl; return false; }
      return true;
    }
    if (dmi_colorize(ctx_, P, points_.data(), DMI_F32, views_, colors_.data(), K_.data(), RT_.data(), W_, H_,
                     MeanColoration.data(), MedianColoration.data(), NbProjectedDepthMap.data()) != DMI_OK)
    { std::cerr << dmi_last_error(ctx_) << std::endl; return false; }
    return true;
  }

  // point-data arrays by the names the reference gives them (MeshColoration.cxx:119,127,133)
  std::vector<uint8_t> MeanColoration, MedianColoration;
  std::vector<int32_t> NbProjectedDepthMap;

private:
  dmi_ctx* ctx_ = nullptr;
  dmi_group* group_ = nullptr;
  bool hasInput_ = false;
  int views_ = 0, W_ = 0, H_ = 0;
  std::vector<float> points_;
  std::vector<uint8_t> colors_;
  std::vector<double> K_, RT_;
};

}  // namespace dmihost
