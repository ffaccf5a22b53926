// adapters/DmiMeshColoration.cxx -- VTK-side adapter of the coloration seam.
//
// A replacement for Coloration/MeshColoration.cxx behind the reference's UNCHANGED header (MeshColoration.h:42-62):
// same constructors, SetInput / GetOutput / ProcessColoration, same three point-data arrays on the output mesh
// ("MeanColoration", "MedianColoration", "NbProjectedDepthMap", MeshColoration.cxx:113-133,194-196).  The per-point
// loop (:140-192) runs on the GPU through dmi_colorize (include/dmi_b200.h).  Build it INSTEAD of
// MeshColoration.cxx (adapters/CMakeLists.txt); Coloration/main.cxx and ReconstructionData.cxx stay untouched.
#include "MeshColoration.h"

#include "dmi_b200.h"

#include "Helper.h"
#include "ReconstructionData.h"

#include "vtkImageData.h"
#include "vtkIntArray.h"
#include "vtkMatrix3x3.h"
#include "vtkMatrix4x4.h"
#include "vtkPointData.h"
#include "vtkPoints.h"
#include "vtkPolyData.h"
#include "vtkUnsignedCharArray.h"

#include <cstdlib>
#include <cstring>
#include <iostream>

MeshColoration::MeshColoration() : OutputMesh(NULL) {}

MeshColoration::MeshColoration(vtkPolyData* mesh, std::string vtiList, std::string krtdList) : OutputMesh(NULL)
{
  this->SetInput(mesh);
  const std::vector<std::string> images = help::ExtractAllFilePath(vtiList.c_str());
  const std::vector<std::string> cameras = help::ExtractAllFilePath(krtdList.c_str());
  if (cameras.size() < images.size())
  {
    std::cerr << "Error, not enough krtd file for each vti file" << std::endl;
    return;
  }
  for (size_t v = 0; v < images.size(); v++) this->DataList.push_back(new ReconstructionData(images[v], cameras[v]));
}

MeshColoration::~MeshColoration()
{
  if (this->OutputMesh) this->OutputMesh->Delete();
  for (size_t v = 0; v < this->DataList.size(); v++) delete this->DataList[v];
}

void MeshColoration::SetInput(vtkPolyData* mesh)
{
  if (this->OutputMesh) this->OutputMesh->Delete();
  this->OutputMesh = vtkPolyData::New();
  this->OutputMesh->DeepCopy(mesh);
}

vtkPolyData* MeshColoration::GetOutput() { return this->OutputMesh; }

bool MeshColoration::ProcessColoration()
{
  const int nViews = (int)this->DataList.size();
  if (this->OutputMesh == NULL || nViews == 0)
  {
    std::cerr << "Error when input has been set or during reading vti/krtd file path" << std::endl;
    return false;
  }
  vtkPoints* points = this->OutputMesh->GetPoints();
  const vtkIdType nPoints = points->GetNumberOfPoints();
  // every view is tested against view 0's dimensions (MeshColoration.cxx:110,158-163)
  const int* dims = this->DataList[0]->GetDepthMap()->GetDimensions();
  const int W = dims[0], H = dims[1];
  const size_t npix = (size_t)W * H;

  std::vector<unsigned char> colors(npix * 3 * (size_t)nViews);
  std::vector<double> K(16 * (size_t)nViews, 0.0), RT(16 * (size_t)nViews, 0.0);
  for (int v = 0; v < nViews; v++)
  {
    ReconstructionData* data = this->DataList[v];
    vtkUnsignedCharArray* c = vtkUnsignedCharArray::SafeDownCast(data->GetDepthMap()->GetPointData()->GetArray("Color"));
    if (!c || (size_t)c->GetNumberOfTuples() != npix || c->GetNumberOfComponents() != 3)
    {
      std::cerr << "Error, no 'Color' array exists" << std::endl;      // ReconstructionData.cxx:97-101
      return false;
    }
    memcpy(&colors[npix * 3 * (size_t)v], c->GetPointer(0), npix * 3);
    for (int r = 0; r < 4; r++)
      for (int q = 0; q < 4; q++)
      {
        K[16 * (size_t)v + 4 * r + q] = data->Get4MatrixK()->GetElement(r, q);
        RT[16 * (size_t)v + 4 * r + q] = data->GetMatrixTR()->GetElement(r, q);
      }
  }

  vtkUnsignedCharArray* mean = vtkUnsignedCharArray::New();
  vtkUnsignedCharArray* median = vtkUnsignedCharArray::New();
  vtkIntArray* count = vtkIntArray::New();
  mean->SetNumberOfComponents(3); mean->SetNumberOfTuples(nPoints); mean->SetName("MeanColoration");
  median->SetNumberOfComponents(3); median->SetNumberOfTuples(nPoints); median->SetName("MedianColoration");
  count->SetNumberOfComponents(1); count->SetNumberOfTuples(nPoints); count->SetName("NbProjectedDepthMap");
  if (nPoints > 0)
  {
    memset(mean->GetPointer(0), 0, (size_t)nPoints * 3);
    memset(median->GetPointer(0), 0, (size_t)nPoints * 3);
    memset(count->GetPointer(0), 0, (size_t)nPoints * sizeof(int));
  }

  bool ok = true;
  if (nPoints > 0)
  {
    vtkDataArray* xyz = points->GetData();
    const int type = xyz->GetDataType();
    std::vector<double> promoted;
    const void* coords = xyz->GetVoidPointer(0);
    int xyzType = type == VTK_FLOAT ? DMI_F32 : DMI_F64;
    if (type != VTK_FLOAT && type != VTK_DOUBLE)
    {
      promoted.resize(3 * (size_t)nPoints);                           // GetPoint promotes any storage to double (:147-148)
      for (vtkIdType p = 0; p < nPoints; p++) points->GetPoint(p, &promoted[3 * (size_t)p]);
      coords = promoted.data();
    }
    dmi_ctx* ctx = NULL;
    const char* dev = getenv("DMI_DEVICE");
    if (dmi_create(dev ? atoi(dev) : 0, &ctx) != DMI_OK)
    {
      std::cerr << "dmi_create: " << dmi_last_error(NULL) << std::endl;
      ok = false;
    }
    else
    {
      if (dmi_colorize(ctx, (size_t)nPoints, coords, xyzType, nViews, colors.data(), K.data(), RT.data(), W, H,
                       mean->GetPointer(0), median->GetPointer(0), count->GetPointer(0)) != DMI_OK)
      {
        std::cerr << "dmi_colorize: " << dmi_last_error(ctx) << std::endl;
        ok = false;
      }
      dmi_destroy(ctx);
    }
  }
  if (ok)
  {
    vtkPointData* pd = this->OutputMesh->GetPointData();
    pd->AddArray(mean);
    pd->AddArray(median);
    pd->AddArray(count);
  }
  mean->Delete(); median->Delete(); count->Delete();
  return ok;
}
