// adapters/vtkDmiReconstruction.cxx -- VTK-side adapter of the TSDF seam.
//
// Defines the two free functions that the reference's filter forward-declares and calls
// (Reconstruction/vtkCudaReconstructionFilter.cxx:65-71, called at :171-176) with their EXACT signatures, on top of
// the C ABI of libdmi_b200.so (include/dmi_b200.h).  Build it INSTEAD of Reconstruction/CudaReconstruction.cu
// (adapters/CMakeLists.txt): vtkCudaReconstructionFilter.cxx, main.cxx and ReconstructionData.cxx stay untouched.
//
//   CudaInitialize      -> dmi_initialize           (CudaReconstruction.cu:269-298)
//   ProcessDepthMap<T>  -> dmi_volume_begin / dmi_volume_integrate_host per batch of files / dmi_volume_end
//                                                    (CudaReconstruction.cu:302-386)
// Kept from the reference: the call accumulates ONTO io_scalar (:323-327), views are integrated in list order
// (:343), the best-cost threshold is applied to every view (:348; here on the GPU, fused into the view
// preparation), false + message on empty lists (:308-312), progress on stdout (:345,383).  Changed on purpose: CUDA
// errors are reported and make the call return false instead of exit()ing the process (:68-76).
// Like the reference (process-global __constant__ state, :55-63) the adapter keeps ONE context per process.
#include "dmi_b200.h"

#include "ReconstructionData.h"

#include "vtkDoubleArray.h"
#include "vtkImageData.h"
#include "vtkMatrix4x4.h"
#include "vtkPointData.h"

#include <algorithm>
#include <cstdlib>
#include <iostream>
#include <string>
#include <vector>

namespace
{
dmi_ctx* g_ctx = NULL;

dmi_ctx* Context()
{
  if (!g_ctx)
  {
    const char* dev = getenv("DMI_DEVICE");
    if (dmi_create(dev ? atoi(dev) : 0, &g_ctx) != DMI_OK)
    {
      std::cerr << "dmi_create: " << dmi_last_error(NULL) << std::endl;
      g_ctx = NULL;
    }
  }
  return g_ctx;
}

bool Check(int rc, const char* what)
{
  if (rc == DMI_OK) return true;
  std::cerr << what << ": " << dmi_last_error(g_ctx) << std::endl;
  return false;
}

void Flatten(vtkMatrix4x4* m, double out[16])
{
  for (int r = 0; r < 4; r++)
    for (int c = 0; c < 4; c++) out[4 * r + c] = m->GetElement(r, c);
}
}  // namespace

void CudaInitialize(vtkMatrix4x4* i_gridMatrix, int h_gridDims[3], double h_gridOrig[3], double h_gridSpacing[3],
                    double h_rayPThick, double h_rayPRho, double h_rayPEta, double h_rayPDelta, int h_depthMapDim[2])
{
  if (!Context()) return;
  double gm[16];
  Flatten(i_gridMatrix, gm);
  Check(dmi_initialize(g_ctx, gm, h_gridDims, h_gridOrig, h_gridSpacing, h_rayPThick, h_rayPRho, h_rayPEta, h_rayPDelta,
                       h_depthMapDim), "CudaInitialize");
}

template <typename TVolumetric>
bool ProcessDepthMap(std::vector<std::string> vtiList, std::vector<std::string> krtdList, double thresholdBestCost,
                     vtkDoubleArray* io_scalar)
{
  if (vtiList.size() == 0 || krtdList.size() == 0)
  {
    std::cerr << "Error, no depthMap or KRTD matrix have been loaded" << std::endl;
    return false;
  }
  if (!Context()) return false;
  const bool isDouble = sizeof(TVolumetric) == sizeof(double);
  const size_t nbVoxels = (size_t)io_scalar->GetNumberOfTuples();
  const int nbDepthMap = (int)vtiList.size();
  std::cout << "START CUDA ON " << nbDepthMap << " Depth map" << std::endl;

  // ProcessDepthMap<float> converts io_scalar to float first and back at the end (:323-324, :371)
  std::vector<TVolumetric> narrow;
  void* volume = io_scalar->GetPointer(0);
  if (!isDouble)
  {
    narrow.resize(nbVoxels);
    for (size_t i = 0; i < nbVoxels; i++) narrow[i] = (TVolumetric)io_scalar->GetValue((vtkIdType)i);
    volume = narrow.data();
  }
  if (!Check(dmi_volume_begin(g_ctx, volume, isDouble ? DMI_F64 : DMI_F32), "ProcessDepthMap")) return false;

  // files are read view by view like the reference's loop; a batch of them goes to the GPU in one call
  size_t npix = 0, batch = 1;
  std::vector<double> depths, costs, K, RT;
  int done = 0;
  while (done < nbDepthMap)
  {
    depths.clear(); costs.clear(); K.clear(); RT.clear();
    bool haveCost = true;
    int n = 0;
    for (; done + n < nbDepthMap && (size_t)n < batch; n++)
    {
      std::cout << "\r" << (100 * (done + n)) / nbDepthMap << " %" << std::flush;
      ReconstructionData data(vtiList[done + n], krtdList[done + n]);
      vtkPointData* pd = data.GetDepthMap()->GetPointData();
      vtkDoubleArray* d = vtkDoubleArray::SafeDownCast(pd->GetArray("Depths"));
      vtkDoubleArray* c = vtkDoubleArray::SafeDownCast(pd->GetArray("Best Cost Values"));
      if (!d) { std::cerr << "Error, no 'Depths' array in " << vtiList[done + n] << std::endl; return false; }
      const size_t count = (size_t)d->GetNumberOfTuples();
      if (npix == 0)
      {
        npix = count;
        batch = std::max<size_t>(1, std::min<size_t>(32, (256u << 20) / (npix * sizeof(double))));
      }
      if (count != npix) { std::cerr << "Error, depth maps of different sizes" << std::endl; return false; }
      depths.insert(depths.end(), d->GetPointer(0), d->GetPointer(0) + npix);
      // ApplyDepthThresholdFilter leaves the map alone when the two arrays differ in size (ReconstructionData.cxx:156-157)
      if (c && (size_t)c->GetNumberOfTuples() == npix) costs.insert(costs.end(), c->GetPointer(0), c->GetPointer(0) + npix);
      else haveCost = false;
      double m[16];
      Flatten(data.Get4MatrixK(), m);
      K.insert(K.end(), m, m + 16);
      Flatten(data.GetMatrixTR(), m);
      RT.insert(RT.end(), m, m + 16);
    }
    if (!haveCost && !costs.empty())
    {
      // mixed batch: integrate its views one by one so that each gets its own filter decision
      size_t cpos = 0;
      for (int v = 0; v < n; v++)
      {
        ReconstructionData data(vtiList[done + v], krtdList[done + v]);
        vtkDoubleArray* c = vtkDoubleArray::SafeDownCast(data.GetDepthMap()->GetPointData()->GetArray("Best Cost Values"));
        const bool has = c && (size_t)c->GetNumberOfTuples() == npix;
        if (!Check(dmi_volume_integrate_host(g_ctx, 1, &depths[npix * v], has ? &costs[npix * cpos] : NULL, thresholdBestCost,
                                             &K[16 * v], &RT[16 * v]), "ProcessDepthMap")) return false;
        if (has) cpos++;
      }
    }
    else if (!Check(dmi_volume_integrate_host(g_ctx, n, depths.data(), haveCost ? costs.data() : NULL, thresholdBestCost,
                                              K.data(), RT.data()), "ProcessDepthMap")) return false;
    done += n;
  }
  if (!Check(dmi_volume_end(g_ctx, volume), "ProcessDepthMap")) return false;
  if (!isDouble)
    for (size_t i = 0; i < nbVoxels; i++) io_scalar->SetValue((vtkIdType)i, (double)narrow[i]);
  std::cout << "\r" << "100 %" << std::flush << std::endl << std::endl;
  return true;
}

// the two instantiations the reference provides (CudaReconstruction.cu:390-400)
template bool ProcessDepthMap<float>(std::vector<std::string> vtiList, std::vector<std::string> krtdList,
                                     double thresholdBestCost, vtkDoubleArray* io_scalar);
template bool ProcessDepthMap<double>(std::vector<std::string> vtiList, std::vector<std::string> krtdList,
                                      double thresholdBestCost, vtkDoubleArray* io_scalar);
