/*
 * oracle/adapter_harness.cpp -- TEST INFRASTRUCTURE ONLY.
 *
 * Drives the reference's OWN operator classes, compiled unmodified from /root/reference against the VTK stand-in of
 * oracle/vtk_shim/ (oracle/Makefile):
 *     vtkCudaReconstructionFilter (Reconstruction/vtkCudaReconstructionFilter.{h,cxx}): SetInputData, the Set* macros,
 *                                 SetGridMatrix, Update(), output cell array "reconstruction_scalar" (.cxx:96-151)
 *     MeshColoration              (Coloration/MeshColoration.h:42-62)
 * through their public interfaces.  The same file is linked twice:
 *   oracle/_ref/libadapter_vtk.so  with adapters/vtkDmiReconstruction.cxx + adapters/DmiMeshColoration.cxx + libdmi_b200.so
 *                                  = the drop-in: the reference's filter class calling THIS repository's GPU path
 *   oracle/_ref/libref_full.so     with the reference's Reconstruction/CudaReconstruction.cu (nvcc, sm_100a) and
 *                                  Coloration/MeshColoration.cxx = the reference end to end (baseline B1 as shipped)
 * Views are handed over like the reference receives them: list files and .krtd files on disk, the .vti images through
 * the stand-in reader's registry.
 */
#include "vtkCellData.h"
#include "vtkDoubleArray.h"
#include "vtkImageData.h"
#include "vtkIntArray.h"
#include "vtkMatrix4x4.h"
#include "vtkNew.h"
#include "vtkPointData.h"
#include "vtkPoints.h"
#include "vtkPolyData.h"
#include "vtkUnsignedCharArray.h"

#include "MeshColoration.h"
#include "vtkCudaReconstructionFilter.h"

#include <cstdint>
#include <cstdio>
#include <string>

static bool write_text(const std::string& path, const std::string& text)
{
  FILE* f = fopen(path.c_str(), "w");
  if (!f) return false;
  const bool ok = fwrite(text.data(), 1, text.size(), f) == text.size();
  return fclose(f) == 0 && ok;
}

static std::string num(double v)
{
  char b[64];
  snprintf(b, sizeof(b), "%.17g", v);
  return b;
}

/* list files, .krtd files and registered images for nViews views; any of depths / bestCost / colors may be NULL */
static bool stage_views(const std::string& dir, int nViews, const double* depths, const double* bestCost,
                        const uint8_t* colors, const double* K, const double* RT, int W, int H)
{
  std::string vtiList, krtdList;
  const size_t npix = (size_t)W * H;
  for (int v = 0; v < nViews; v++)
  {
    char name[64];
    snprintf(name, sizeof(name), "view_%05d", v);
    vtiList += std::string(name) + ".vti\n";
    krtdList += std::string(name) + ".krtd\n";
    const double* k = K + 16 * (size_t)v;
    const double* rt = RT + 16 * (size_t)v;
    std::string t;
    for (int r = 0; r < 3; r++) t += num(k[4 * r]) + " " + num(k[4 * r + 1]) + " " + num(k[4 * r + 2]) + "\n";
    t += "\n";
    for (int r = 0; r < 3; r++) t += num(rt[4 * r]) + " " + num(rt[4 * r + 1]) + " " + num(rt[4 * r + 2]) + "\n";
    t += "\n";
    t += num(rt[3]) + " " + num(rt[7]) + " " + num(rt[11]) + "\n";
    if (!write_text(dir + "/" + name + ".krtd", t)) return false;
    vtkImageData* img = vtkImageData::New();
    img->SetDimensions(W, H, 1);
    const double* src[2] = {depths, bestCost};
    const char* names[2] = {"Depths", "Best Cost Values"};
    for (int a = 0; a < 2; a++)
      if (src[a])
      {
        vtkDoubleArray* arr = vtkDoubleArray::New();
        arr->SetName(names[a]);
        arr->SetNumberOfComponents(1);
        arr->SetNumberOfTuples((vtkIdType)npix);
        memcpy(arr->GetPointer(0), src[a] + npix * (size_t)v, npix * sizeof(double));
        img->GetPointData()->AddArray(arr);
        arr->Delete();
      }
    if (colors)
    {
      vtkUnsignedCharArray* c = vtkUnsignedCharArray::New();
      c->SetName("Color");
      c->SetNumberOfComponents(3);
      c->SetNumberOfTuples((vtkIdType)npix);
      memcpy(c->GetPointer(0), colors + npix * 3 * (size_t)v, npix * 3);
      img->GetPointData()->AddArray(c);
      c->Delete();
    }
    vtkStandIn::RegisterImage(dir + "/" + name + ".vti", img);
    img->Delete();
  }
  return write_text(dir + "/vtiList.txt", vtiList) && write_text(dir + "/kList.txt", krtdList);
}

/* vtkCudaReconstructionFilter end to end; out receives the (dims-1)^3 cell scalars.  Returns 0 on success,
 * 1 when RequestData reported failure, 2 on I/O problems, 3 when the output array is missing.
 * execSeconds (optional) receives the filter's ExecutionTime. */
extern "C" int harness_filter_run(const char* workdir, const double* gridMatrix, const int* dims, const double* origin,
                                  const double* spacing, double thick, double rho, double eta, double delta,
                                  double threshold, int nViews, const double* depths, const double* bestCost,
                                  const double* K, const double* RT, int W, int H, double* out, double* execSeconds)
{
  const std::string dir(workdir);
  if (!stage_views(dir, nViews, depths, bestCost, NULL, K, RT, W, H)) return 2;
  vtkNew<vtkImageData> grid;
  grid->SetDimensions(dims[0], dims[1], dims[2]);
  grid->SetOrigin(origin[0], origin[1], origin[2]);
  grid->SetSpacing(spacing[0], spacing[1], spacing[2]);
  vtkNew<vtkMatrix4x4> gm;
  for (int r = 0; r < 4; r++)
    for (int c = 0; c < 4; c++) gm->SetElement(r, c, gridMatrix[4 * r + c]);
  int rc = 0;
  vtkCudaReconstructionFilter* filter = vtkCudaReconstructionFilter::New();
  filter->SetInputData(grid.Get());
  filter->SetFilePathKRTD((dir + "/kList.txt").c_str());
  filter->SetFilePathVTI((dir + "/vtiList.txt").c_str());
  filter->SetGridMatrix(gm.Get());
  filter->SetRayPotentialRho(rho);
  filter->SetRayPotentialThickness(thick);
  filter->SetRayPotentialEta(eta);
  filter->SetRayPotentialDelta(delta);
  filter->SetThresholdBestCost(threshold);
  filter->Update();
  if (!filter->GetLastRequestDataStatus()) rc = 1;
  else
  {
    vtkDoubleArray* s = vtkDoubleArray::SafeDownCast(filter->GetOutput()->GetCellData()->GetArray("reconstruction_scalar"));
    if (!s) rc = 3;
    else memcpy(out, s->GetPointer(0), (size_t)s->GetNumberOfTuples() * sizeof(double));
    if (execSeconds) *execSeconds = filter->GetExecutionTime();
  }
  filter->Delete();
  vtkStandIn::ClearImages();
  return rc;
}

/* MeshColoration end to end (same contract as oracle/ref_coloration_harness.cpp) */
extern "C" int harness_coloration_run(const char* workdir, size_t nPoints, const void* xyz, int xyzType, int nViews,
                                      const uint8_t* colors, const double* K, const double* RT, int W, int H,
                                      uint8_t* mean, uint8_t* median, int32_t* nb)
{
  const std::string dir(workdir);
  if (!stage_views(dir, nViews, NULL, NULL, colors, K, RT, W, H)) return 2;
  vtkPolyData* mesh = vtkPolyData::New();
  vtkPoints* pts = vtkPoints::New();
  if (xyzType == 1) pts->SetDataTypeToDouble();
  pts->SetNumberOfPoints((vtkIdType)nPoints);
  for (size_t p = 0; p < nPoints; p++)
  {
    if (xyzType == 1) { const double* q = (const double*)xyz + 3 * p; pts->SetPoint((vtkIdType)p, q[0], q[1], q[2]); }
    else { const float* q = (const float*)xyz + 3 * p; pts->SetPoint((vtkIdType)p, q[0], q[1], q[2]); }
  }
  mesh->SetPoints(pts);
  pts->Delete();
  int rc = 0;
  {
    MeshColoration coloration(mesh, dir + "/vtiList.txt", dir + "/kList.txt");
    if (!coloration.ProcessColoration()) rc = 1;
    else
    {
      vtkPointData* pd = coloration.GetOutput()->GetPointData();
      vtkUnsignedCharArray* m = vtkUnsignedCharArray::SafeDownCast(pd->GetArray("MeanColoration"));
      vtkUnsignedCharArray* d = vtkUnsignedCharArray::SafeDownCast(pd->GetArray("MedianColoration"));
      vtkIntArray* n = vtkIntArray::SafeDownCast(pd->GetArray("NbProjectedDepthMap"));
      if (!m || !d || !n) rc = 3;
      else
      {
        memcpy(mean, m->GetPointer(0), nPoints * 3);
        memcpy(median, d->GetPointer(0), nPoints * 3);
        memcpy(nb, n->GetPointer(0), nPoints * sizeof(int32_t));
      }
    }
  }
  mesh->Delete();
  vtkStandIn::ClearImages();
  return rc;
}
