/*
 * oracle/ref_coloration_harness.cpp -- TEST INFRASTRUCTURE ONLY.
 *
 * Drives the reference's OWN coloration code -- Coloration/MeshColoration.cxx, Sources/ReconstructionData.cxx and
 * Sources/Helper.h, compiled unmodified from /root/reference by oracle/Makefile against the VTK stand-in under
 * oracle/vtk_shim/ -- through its public interface (MeshColoration.h:45-56):
 *     MeshColoration(vtkPolyData*, vtiListFile, krtdListFile); ProcessColoration(); GetOutput()
 * The views are handed over the way the reference receives them: list files and .krtd files on disk (written
 * here with 17 significant digits, which round-trips every finite double through help::ReadKrtdFile), and the
 * "Color" images through the stand-in's vtkXMLImageDataReader registry (the .vti XML itself is not the path
 * under test).  Output -> oracle/_ref/libref_coloration.so, used to pin oracle/color_oracle.c bit for bit.
 */
#include "vtkImageData.h"
#include "vtkPointData.h"
#include "vtkPoints.h"
#include "vtkPolyData.h"
#include "vtkUnsignedCharArray.h"
#include "vtkIntArray.h"

#include "MeshColoration.h"

#include <cstdio>
#include <cstdint>
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

/* K, RT: [nViews][16] row-major 4x4 (only K's 3x3 and RT's 3x4 are stored in a .krtd, Helper.h:105-168).
 * xyzType: 0 = float32 points (vtkPoints' default storage), 1 = float64.  Returns 0 on success. */
extern "C" int ref_coloration_run(const char* workdir, size_t nPoints, const void* xyz, int xyzType, int nViews,
                                  const uint8_t* colors, const double* K, const double* RT, int W, int H,
                                  uint8_t* mean, uint8_t* median, int32_t* nb)
{
  const std::string dir(workdir);
  std::string vtiList, krtdList;
  const size_t npix = (size_t)W * H;
  for (int v = 0; v < nViews; v++)
  {
    char name[64];
    snprintf(name, sizeof(name), "view_%05d", v);
    /* a leading token before the file name: ExtractAllFilePath keeps the LAST space-separated token (Helper.h:86-97) */
    vtiList += std::string("frame") + " " + name + ".vti\n";
    krtdList += std::string(name) + ".krtd\n";
    const double* k = K + 16 * (size_t)v;
    const double* rt = RT + 16 * (size_t)v;
    std::string t;
    for (int r = 0; r < 3; r++) t += num(k[4 * r]) + " " + num(k[4 * r + 1]) + " " + num(k[4 * r + 2]) + "\n";
    t += "\n";
    for (int r = 0; r < 3; r++) t += num(rt[4 * r]) + " " + num(rt[4 * r + 1]) + " " + num(rt[4 * r + 2]) + "\n";
    t += "\n";
    t += num(rt[3]) + " " + num(rt[7]) + " " + num(rt[11]) + "\n";
    if (!write_text(dir + "/" + name + ".krtd", t)) return 2;

    vtkImageData* img = vtkImageData::New();
    img->SetDimensions(W, H, 1);
    vtkUnsignedCharArray* c = vtkUnsignedCharArray::New();
    c->SetName("Color");
    c->SetNumberOfComponents(3);
    c->SetNumberOfTuples((vtkIdType)npix);
    memcpy(c->GetPointer(0), colors + npix * 3 * (size_t)v, npix * 3);
    img->GetPointData()->AddArray(c);
    c->Delete();
    vtkStandIn::RegisterImage(dir + "/" + name + ".vti", img);
    img->Delete();
  }
  if (!write_text(dir + "/vtiList.txt", vtiList) || !write_text(dir + "/kList.txt", krtdList)) return 2;

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

/* ReconstructionData::TransformWorldToDepthMapPosition alone (ReconstructionData.cxx:169-182), for known answers. */
#include "ReconstructionData.h"
extern "C" int ref_world_to_pixel(const char* workdir, const double* K16, const double* RT16, const double* p, int* pixel)
{
  const uint8_t one[3] = {0, 0, 0};
  const std::string dir(workdir);
  std::string t;
  for (int r = 0; r < 3; r++) t += num(K16[4 * r]) + " " + num(K16[4 * r + 1]) + " " + num(K16[4 * r + 2]) + "\n";
  t += "\n";
  for (int r = 0; r < 3; r++) t += num(RT16[4 * r]) + " " + num(RT16[4 * r + 1]) + " " + num(RT16[4 * r + 2]) + "\n";
  t += "\n";
  t += num(RT16[3]) + " " + num(RT16[7]) + " " + num(RT16[11]) + "\n";
  if (!write_text(dir + "/one.krtd", t)) return 2;
  vtkImageData* img = vtkImageData::New();
  img->SetDimensions(1, 1, 1);
  vtkUnsignedCharArray* c = vtkUnsignedCharArray::New();
  c->SetName("Color");
  c->SetNumberOfComponents(3);
  c->SetNumberOfTuples(1);
  memcpy(c->GetPointer(0), one, 3);
  img->GetPointData()->AddArray(c);
  c->Delete();
  vtkStandIn::RegisterImage(dir + "/one.vti", img);
  img->Delete();
  {
    ReconstructionData data(dir + "/one.vti", dir + "/one.krtd");
    data.TransformWorldToDepthMapPosition(p, pixel);
  }
  vtkStandIn::ClearImages();
  return 0;
}
