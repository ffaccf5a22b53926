/*
 * oracle/vtk_shim/vtk_standin.h -- TEST INFRASTRUCTURE ONLY.
 *
 * A minimal stand-in for the handful of VTK 6-8 classes the reference touches on its hot path, so that the
 * reference's OWN translation units
 *     Coloration/MeshColoration.cxx, Sources/ReconstructionData.cxx, Sources/Helper.h,
 *     Reconstruction/CudaReconstruction.cu, Reconstruction/vtkCudaReconstructionFilter.cxx
 * compile UNMODIFIED, from where they lie under /root/reference, in an image without VTK (oracle/Makefile).
 * VTK is an un-vendored, version-unpinned dependency of the reference (CMakeLists.txt:8-18: pre-9 component
 * names => VTK 6-8); what is restated here is the PUBLISHED behaviour of those classes, in particular the
 * arithmetic that sits on the coloration path:
 *   vtkMatrix4x4::Multiply4x4            a[i][0]*b[0][k] + a[i][1]*b[1][k] + a[i][2]*b[2][k] + a[i][3]*b[3][k]
 *   vtkTransform::SetMatrix(M)           concatenation := identity, then Concatenate(M) (pre-multiply):
 *                                        Matrix = Multiply4x4(identity, Multiply4x4(identity, M)) on Update()
 *   vtkLinearTransform::TransformPoint   m[r][0]*x + m[r][1]*y + m[r][2]*z + m[r][3], in double, left to right
 *   vtkLinearTransform::TransformVector  m[r][0]*x + m[r][1]*y + m[r][2]*z
 *   vtkImageData::ComputePointId         (k - e4)*d0*d1 + (j - e2)*d0 + (i - e0)
 *   vtkDataArray::SetTuple1/3(double)    static_cast<ValueType>   (truncation for integer arrays)
 *   vtkDataArray::GetTuple1/3            static_cast<double>
 *   vtkPoints                            float32 storage by default, GetPoint promotes to double
 * vtkXMLImageDataReader does not parse XML here: Update() hands out a fresh DEEP copy of the image that a test
 * registered under that file name (vtkStandIn::RegisterImage), which is what reading the file again would give.
 *
 * Nothing under cudadepthmapintegration_b200/ may include this header; adapters/ is compile-CHECKED against it.
 */
#ifndef DMI_VTK_STANDIN_H
#define DMI_VTK_STANDIN_H

#include <cstddef>
#include <cstring>
#include <fstream>
#include <iostream>
#include <map>
#include <sstream>
#include <string>
#include <vector>
#include <unistd.h>

typedef long long vtkIdType;
#define VTK_UNSIGNED_CHAR 3
#define VTK_INT 6
#define VTK_FLOAT 10
#define VTK_DOUBLE 11
using std::ostream;
using std::cerr;
using std::cout;
using std::endl;

#define vtkNotUsed(x)
#define vtkTypeMacro(thisClass, superClass) \
  typedef superClass Superclass;            \
  const char* GetClassName() const { return #thisClass; }
#define vtkStandardNewMacro(thisClass) \
  thisClass* thisClass::New() { return new thisClass; }
#define vtkSetMacro(name, type) \
  virtual void Set##name(type _arg) { if (this->name != _arg) { this->name = _arg; this->Modified(); } }
#define vtkGetMacro(name, type) \
  virtual type Get##name() { return this->name; }
/* the reference declares `const char* FilePathVTI` and uses vtkSetStringMacro on it: VTK's macro allocates a
 * copy with new[] and assigns it (legal for a const char* member as long as nothing writes through it) */
#define vtkSetStringMacro(name)                                                                      \
  virtual void Set##name(const char* _arg)                                                           \
  {                                                                                                  \
    if (this->name == NULL && _arg == NULL) return;                                                  \
    if (this->name && _arg && (!strcmp(this->name, _arg))) return;                                   \
    delete[] this->name;                                                                             \
    if (_arg) { size_t n = strlen(_arg) + 1; char* cp1 = new char[n]; memcpy(cp1, _arg, n); this->name = cp1; } \
    else this->name = NULL;                                                                          \
    this->Modified();                                                                                \
  }
#define vtkSetObjectImplementationMacro(thisClass, name, type)                                       \
  void thisClass::Set##name(type* _arg)                                                              \
  {                                                                                                  \
    if (this->name != _arg)                                                                          \
    {                                                                                                \
      type* tempSGMacroVar = this->name;                                                             \
      this->name = _arg;                                                                             \
      if (this->name != NULL) this->name->Register(this);                                            \
      if (tempSGMacroVar != NULL) tempSGMacroVar->UnRegister(this);                                  \
      this->Modified();                                                                              \
    }                                                                                                \
  }
#define vtkErrorMacro(x) do { std::cerr << "ERROR: " x << std::endl; } while (0)

class vtkIndent {};

class vtkObjectBase
{
public:
  vtkObjectBase() : ReferenceCount(1) {}
  virtual ~vtkObjectBase() {}
  virtual void Delete() { this->UnRegister(0); }
  virtual void Register(vtkObjectBase*) { ++this->ReferenceCount; }
  virtual void UnRegister(vtkObjectBase*) { if (--this->ReferenceCount <= 0) delete this; }
  int GetReferenceCount() const { return this->ReferenceCount; }
  virtual void Modified() {}
  virtual void PrintSelf(ostream&, vtkIndent) {}
protected:
  int ReferenceCount;
};
typedef vtkObjectBase vtkObject;

template <class T>
class vtkSmartPointer
{
public:
  vtkSmartPointer() : P(0) {}
  vtkSmartPointer(T* p) : P(p) { if (P) P->Register(0); }            /* like VTK: takes ANOTHER reference */
  vtkSmartPointer(const vtkSmartPointer& o) : P(o.P) { if (P) P->Register(0); }
  ~vtkSmartPointer() { if (P) P->UnRegister(0); }
  vtkSmartPointer& operator=(const vtkSmartPointer& o) { if (o.P) o.P->Register(0); if (P) P->UnRegister(0); P = o.P; return *this; }
  static vtkSmartPointer New() { vtkSmartPointer s; s.P = T::New(); return s; }
  T* operator->() const { return P; }
  T* Get() const { return P; }
  T* GetPointer() const { return P; }
  operator T*() const { return P; }
private:
  T* P;
};

template <class T>
class vtkNew
{
public:
  vtkNew() : P(T::New()) {}
  ~vtkNew() { P->Delete(); }
  T* operator->() const { return P; }
  T* Get() const { return P; }
  T* GetPointer() const { return P; }
  operator T*() const { return P; }
private:
  vtkNew(const vtkNew&);
  void operator=(const vtkNew&);
  T* P;
};

/* ---- matrices ---------------------------------------------------------------------------------------- */

class vtkMatrix3x3 : public vtkObject
{
public:
  static vtkMatrix3x3* New() { return new vtkMatrix3x3; }
  vtkMatrix3x3() { Identity(); }
  void Identity() { for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) Element[i][j] = (i == j) ? 1.0 : 0.0; }
  void SetElement(int i, int j, double v) { Element[i][j] = v; }
  double GetElement(int i, int j) const { return Element[i][j]; }
  double Element[3][3];
};

class vtkMatrix4x4 : public vtkObject
{
public:
  static vtkMatrix4x4* New() { return new vtkMatrix4x4; }
  vtkMatrix4x4() { Identity(); }
  void Identity() { for (int i = 0; i < 4; i++) for (int j = 0; j < 4; j++) Element[i][j] = (i == j) ? 1.0 : 0.0; }
  void SetElement(int i, int j, double v) { Element[i][j] = v; }
  double GetElement(int i, int j) const { return Element[i][j]; }
  void DeepCopy(const vtkMatrix4x4* m) { memcpy(Element, m->Element, sizeof(Element)); }
  /* vtkMatrix4x4::Multiply4x4(const double a[16], const double b[16], double c[16]) */
  static void Multiply4x4(const double a[4][4], const double b[4][4], double c[4][4])
  {
    double acc[4][4];
    for (int i = 0; i < 4; i++)
      for (int k = 0; k < 4; k++)
        acc[i][k] = a[i][0] * b[0][k] + a[i][1] * b[1][k] + a[i][2] * b[2][k] + a[i][3] * b[3][k];
    memcpy(c, acc, sizeof(acc));
  }
  double Element[4][4];
};

/* vtkTransform: only SetMatrix + the two linear transforms the reference calls (ReconstructionData.cxx:173,175,
 * 211,220).  SetMatrix(M) = Identity() on the concatenation, then Concatenate(M) in pre-multiply mode; Update()
 * forms Matrix = identity(input) * concatenation. */
class vtkTransform : public vtkObject
{
public:
  static vtkTransform* New() { return new vtkTransform; }
  vtkTransform() { ident(Pre); ident(Matrix); }
  void SetMatrix(vtkMatrix4x4* m)
  {
    double id[4][4];
    ident(id);
    vtkMatrix4x4::Multiply4x4(id, m->Element, Pre);      /* vtkTransformConcatenation::Concatenate, PreMultiply */
    double in[4][4];
    ident(in);                                           /* no Input transform: Matrix starts as identity */
    vtkMatrix4x4::Multiply4x4(in, Pre, Matrix);          /* vtkTransform::InternalUpdate */
  }
  void TransformPoint(const double in[3], double out[3])
  {
    const double x = Matrix[0][0] * in[0] + Matrix[0][1] * in[1] + Matrix[0][2] * in[2] + Matrix[0][3];
    const double y = Matrix[1][0] * in[0] + Matrix[1][1] * in[1] + Matrix[1][2] * in[2] + Matrix[1][3];
    const double z = Matrix[2][0] * in[0] + Matrix[2][1] * in[1] + Matrix[2][2] * in[2] + Matrix[2][3];
    out[0] = x; out[1] = y; out[2] = z;
  }
  void TransformVector(const double in[3], double out[3])
  {
    const double x = Matrix[0][0] * in[0] + Matrix[0][1] * in[1] + Matrix[0][2] * in[2];
    const double y = Matrix[1][0] * in[0] + Matrix[1][1] * in[1] + Matrix[1][2] * in[2];
    const double z = Matrix[2][0] * in[0] + Matrix[2][1] * in[1] + Matrix[2][2] * in[2];
    out[0] = x; out[1] = y; out[2] = z;
  }
private:
  static void ident(double m[4][4]) { for (int i = 0; i < 4; i++) for (int j = 0; j < 4; j++) m[i][j] = (i == j) ? 1.0 : 0.0; }
  double Pre[4][4], Matrix[4][4];
};

/* ---- data arrays ------------------------------------------------------------------------------------- */

class vtkDataArray : public vtkObject
{
public:
  vtkDataArray() : NumberOfComponents(1) { Legacy[0] = Legacy[1] = Legacy[2] = Legacy[3] = 0; }
  void SetName(const char* n) { Name = n ? n : ""; }
  const char* GetName() const { return Name.c_str(); }
  void SetNumberOfComponents(int n) { NumberOfComponents = n; }
  int GetNumberOfComponents() const { return NumberOfComponents; }
  virtual void SetNumberOfTuples(vtkIdType n) = 0;
  virtual vtkIdType GetNumberOfTuples() const = 0;
  virtual double GetComponent(vtkIdType t, int c) const = 0;
  virtual void SetComponent(vtkIdType t, int c, double v) = 0;
  virtual vtkDataArray* NewInstance() const = 0;
  virtual void DeepCopy(const vtkDataArray* o) = 0;
  virtual int GetDataType() const = 0;
  virtual void* GetVoidPointer(vtkIdType id) = 0;
  void FillComponent(int c, double v) { for (vtkIdType t = 0, n = GetNumberOfTuples(); t < n; t++) SetComponent(t, c, v); }
  double GetTuple1(vtkIdType t) const { return GetComponent(t, 0); }
  void SetTuple1(vtkIdType t, double v) { SetComponent(t, 0, v); }
  void SetTuple3(vtkIdType t, double a, double b, double c) { SetComponent(t, 0, a); SetComponent(t, 1, b); SetComponent(t, 2, c); }
  double* GetTuple3(vtkIdType t) { for (int c = 0; c < 3; c++) Legacy[c] = GetComponent(t, c); return Legacy; }
  void GetTuple(vtkIdType t, double* out) const { for (int c = 0; c < NumberOfComponents; c++) out[c] = GetComponent(t, c); }
protected:
  std::string Name;
  int NumberOfComponents;
  double Legacy[4];
};

template <class V> struct vtkStandInTypeCode;
template <> struct vtkStandInTypeCode<double> { enum { value = VTK_DOUBLE }; };
template <> struct vtkStandInTypeCode<float> { enum { value = VTK_FLOAT }; };
template <> struct vtkStandInTypeCode<unsigned char> { enum { value = VTK_UNSIGNED_CHAR }; };
template <> struct vtkStandInTypeCode<int> { enum { value = VTK_INT }; };

template <class V, class Self>
class vtkStandInArray : public vtkDataArray
{
public:
  int GetDataType() const { return vtkStandInTypeCode<V>::value; }
  void* GetVoidPointer(vtkIdType id) { return Data.data() + id; }
  static Self* New() { return new Self; }
  static Self* SafeDownCast(vtkObjectBase* o) { return dynamic_cast<Self*>(o); }
  void SetNumberOfTuples(vtkIdType n) { Data.resize((size_t)n * NumberOfComponents); }
  vtkIdType GetNumberOfTuples() const { return (vtkIdType)(Data.size() / (size_t)NumberOfComponents); }
  double GetComponent(vtkIdType t, int c) const { return static_cast<double>(Data[(size_t)t * NumberOfComponents + c]); }
  void SetComponent(vtkIdType t, int c, double v) { Data[(size_t)t * NumberOfComponents + c] = static_cast<V>(v); }
  V* GetPointer(vtkIdType id) { return Data.data() + id; }
  V GetValue(vtkIdType id) const { return Data[(size_t)id]; }
  void SetValue(vtkIdType id, V v) { Data[(size_t)id] = v; }
  vtkDataArray* NewInstance() const { return new Self; }
  void DeepCopy(const vtkDataArray* o)
  {
    const Self* s = dynamic_cast<const Self*>(o);
    Name = s->Name; NumberOfComponents = s->NumberOfComponents; Data = s->Data;
  }
  std::vector<V> Data;
};
class vtkDoubleArray : public vtkStandInArray<double, vtkDoubleArray> {};
class vtkFloatArray : public vtkStandInArray<float, vtkFloatArray> {};
class vtkUnsignedCharArray : public vtkStandInArray<unsigned char, vtkUnsignedCharArray> {};
class vtkIntArray : public vtkStandInArray<int, vtkIntArray> {};

/* vtkFieldData::AddArray replaces an array of the same name; GetArray(name) returns NULL when absent */
class vtkFieldData : public vtkObject
{
public:
  ~vtkFieldData() { Clear(); }
  int AddArray(vtkDataArray* a)
  {
    a->Register(this);
    for (size_t i = 0; i < Arrays.size(); i++)
      if (!strcmp(Arrays[i]->GetName(), a->GetName())) { Arrays[i]->UnRegister(this); Arrays[i] = a; return (int)i; }
    Arrays.push_back(a);
    return (int)Arrays.size() - 1;
  }
  vtkDataArray* GetArray(const char* name)
  {
    for (size_t i = 0; i < Arrays.size(); i++) if (!strcmp(Arrays[i]->GetName(), name)) return Arrays[i];
    return 0;
  }
  vtkDataArray* GetArray(int i) { return (i >= 0 && (size_t)i < Arrays.size()) ? Arrays[(size_t)i] : 0; }
  int GetNumberOfArrays() const { return (int)Arrays.size(); }
  void ShallowCopy(vtkFieldData* o) { if (o == this) return; Clear(); for (size_t i = 0; i < o->Arrays.size(); i++) AddArray(o->Arrays[i]); }
  void DeepCopy(vtkFieldData* o)
  {
    if (o == this) return;
    Clear();
    for (size_t i = 0; i < o->Arrays.size(); i++)
    {
      vtkDataArray* c = o->Arrays[i]->NewInstance();
      c->DeepCopy(o->Arrays[i]);
      AddArray(c);
      c->Delete();
    }
  }
protected:
  void Clear() { for (size_t i = 0; i < Arrays.size(); i++) Arrays[i]->UnRegister(this); Arrays.clear(); }
  std::vector<vtkDataArray*> Arrays;
};
class vtkPointData : public vtkFieldData { public: static vtkPointData* New() { return new vtkPointData; } };
class vtkCellData : public vtkFieldData { public: static vtkCellData* New() { return new vtkCellData; } };

/* ---- data sets ----------------------------------------------------------------------------------------- */

class vtkInformationDataObjectKey {};
class vtkDataObject : public vtkObject
{
public:
  static vtkInformationDataObjectKey* DATA_OBJECT() { static vtkInformationDataObjectKey k; return &k; }
};

class vtkDataSet : public vtkDataObject
{
public:
  vtkDataSet() : PointData(new vtkPointData), CellData(new vtkCellData) {}
  ~vtkDataSet() { PointData->Delete(); CellData->Delete(); }
  vtkPointData* GetPointData() { return PointData; }
  vtkCellData* GetCellData() { return CellData; }
protected:
  vtkPointData* PointData;
  vtkCellData* CellData;
};

class vtkImageData : public vtkDataSet
{
public:
  static vtkImageData* New() { return new vtkImageData; }
  static vtkImageData* SafeDownCast(vtkObjectBase* o) { return dynamic_cast<vtkImageData*>(o); }
  vtkImageData()
  {
    for (int a = 0; a < 3; a++) { Dimensions[a] = 0; Origin[a] = 0.0; Spacing[a] = 1.0; Extent[2 * a] = 0; Extent[2 * a + 1] = -1; }
  }
  void SetDimensions(int i, int j, int k) { Dimensions[0] = i; Dimensions[1] = j; Dimensions[2] = k; Extent[0] = Extent[2] = Extent[4] = 0; Extent[1] = i - 1; Extent[3] = j - 1; Extent[5] = k - 1; }
  void SetDimensions(const int d[3]) { SetDimensions(d[0], d[1], d[2]); }
  int* GetDimensions() { return Dimensions; }
  void GetDimensions(int d[3]) { for (int a = 0; a < 3; a++) d[a] = Dimensions[a]; }
  int* GetExtent() { return Extent; }
  void SetOrigin(double x, double y, double z) { Origin[0] = x; Origin[1] = y; Origin[2] = z; }
  void SetOrigin(const double o[3]) { SetOrigin(o[0], o[1], o[2]); }
  void GetOrigin(double o[3]) { for (int a = 0; a < 3; a++) o[a] = Origin[a]; }
  void SetSpacing(double x, double y, double z) { Spacing[0] = x; Spacing[1] = y; Spacing[2] = z; }
  void SetSpacing(const double s[3]) { SetSpacing(s[0], s[1], s[2]); }
  void GetSpacing(double s[3]) { for (int a = 0; a < 3; a++) s[a] = Spacing[a]; }
  vtkIdType GetNumberOfPoints() { return (vtkIdType)Dimensions[0] * Dimensions[1] * Dimensions[2]; }
  vtkIdType GetNumberOfCells()
  {
    vtkIdType n = 1;
    for (int a = 0; a < 3; a++)
    {
      if (Dimensions[a] <= 0) return 0;
      if (Dimensions[a] > 1) n *= Dimensions[a] - 1;
    }
    return n;
  }
  /* vtkStructuredData::ComputePointIdForExtent */
  vtkIdType ComputePointId(int ijk[3])
  {
    const vtkIdType d0 = Extent[1] - Extent[0] + 1, d1 = Extent[3] - Extent[2] + 1;
    return ((vtkIdType)(ijk[2] - Extent[4]) * d1 + (ijk[1] - Extent[2])) * d0 + (ijk[0] - Extent[0]);
  }
  void ShallowCopy(vtkImageData* o) { CopyStructure(o); PointData->ShallowCopy(o->PointData); CellData->ShallowCopy(o->CellData); }
  void DeepCopy(vtkImageData* o) { CopyStructure(o); PointData->DeepCopy(o->PointData); CellData->DeepCopy(o->CellData); }
private:
  void CopyStructure(vtkImageData* o)
  {
    for (int a = 0; a < 3; a++) { Dimensions[a] = o->Dimensions[a]; Origin[a] = o->Origin[a]; Spacing[a] = o->Spacing[a]; }
    for (int a = 0; a < 6; a++) Extent[a] = o->Extent[a];
  }
  int Dimensions[3], Extent[6];
  double Origin[3], Spacing[3];
};

/* vtkPoints: float32 storage unless SetDataTypeToDouble() */
class vtkPoints : public vtkObject
{
public:
  static vtkPoints* New() { return new vtkPoints; }
  vtkPoints() : Data(vtkFloatArray::New()) { Data->SetNumberOfComponents(3); }
  ~vtkPoints() { Data->Delete(); }
  void SetDataTypeToDouble() { Data->Delete(); Data = vtkDoubleArray::New(); Data->SetNumberOfComponents(3); }
  void SetDataTypeToFloat() { Data->Delete(); Data = vtkFloatArray::New(); Data->SetNumberOfComponents(3); }
  void SetNumberOfPoints(vtkIdType n) { Data->SetNumberOfTuples(n); }
  vtkIdType GetNumberOfPoints() const { return Data->GetNumberOfTuples(); }
  void SetPoint(vtkIdType id, double x, double y, double z) { Data->SetTuple3(id, x, y, z); }
  void GetPoint(vtkIdType id, double x[3]) const { Data->GetTuple(id, x); }
  vtkDataArray* GetData() { return Data; }
  void DeepCopy(vtkPoints* o) { Data->Delete(); Data = o->Data->NewInstance(); Data->DeepCopy(o->Data); }
private:
  vtkDataArray* Data;
};

class vtkPolyData : public vtkDataSet
{
public:
  static vtkPolyData* New() { return new vtkPolyData; }
  vtkPolyData() : Points(0) {}
  ~vtkPolyData() { if (Points) Points->UnRegister(this); }
  void SetPoints(vtkPoints* p) { if (p) p->Register(this); if (Points) Points->UnRegister(this); Points = p; }
  vtkPoints* GetPoints() { return Points; }
  void DeepCopy(vtkPolyData* o)
  {
    vtkPoints* p = 0;
    if (o->Points) { p = vtkPoints::New(); p->DeepCopy(o->Points); }
    SetPoints(p);
    if (p) p->Delete();
    PointData->DeepCopy(o->GetPointData());
    CellData->DeepCopy(o->GetCellData());
  }
private:
  vtkPoints* Points;
};

/* ---- "reading" a .vti -------------------------------------------------------------------------------- */

class vtkStandIn
{
public:
  /* The registry keeps its own reference to the image; a later registration under the same name replaces it. */
  static void RegisterImage(const std::string& path, vtkImageData* img)
  {
    std::map<std::string, vtkImageData*>& r = Registry();
    img->Register(0);
    std::map<std::string, vtkImageData*>::iterator it = r.find(path);
    if (it != r.end()) { it->second->UnRegister(0); it->second = img; }
    else r[path] = img;
  }
  static void ClearImages()
  {
    std::map<std::string, vtkImageData*>& r = Registry();
    for (std::map<std::string, vtkImageData*>::iterator it = r.begin(); it != r.end(); ++it) it->second->UnRegister(0);
    r.clear();
  }
  static vtkImageData* Find(const std::string& path)
  {
    std::map<std::string, vtkImageData*>& r = Registry();
    std::map<std::string, vtkImageData*>::iterator it = r.find(path);
    return it == r.end() ? 0 : it->second;
  }
private:
  static std::map<std::string, vtkImageData*>& Registry() { static std::map<std::string, vtkImageData*> r; return r; }
};

class vtkXMLImageDataReader : public vtkObject
{
public:
  static vtkXMLImageDataReader* New() { return new vtkXMLImageDataReader; }
  vtkXMLImageDataReader() : Output(vtkImageData::New()) {}
  ~vtkXMLImageDataReader() { Output->Delete(); }
  void SetFileName(const char* f) { FileName = f ? f : ""; }
  void Update()
  {
    vtkImageData* src = vtkStandIn::Find(FileName);
    if (!src) { std::cerr << "vtk stand-in: no image registered as " << FileName << std::endl; return; }
    Output->DeepCopy(src);          /* every read yields fresh arrays, like parsing the file again */
  }
  vtkImageData* GetOutput() { return Output; }
private:
  std::string FileName;
  vtkImageData* Output;
};

/* ---- the sliver of the pipeline that vtkCudaReconstructionFilter.cxx needs ------------------------------ */

class vtkInformationIntegerVectorKey {};
class vtkInformation : public vtkObject
{
public:
  static vtkInformation* New() { return new vtkInformation; }
  vtkInformation() : Object(0) { for (int a = 0; a < 6; a++) Whole[a] = 0; }
  vtkDataObject* Get(vtkInformationDataObjectKey*) { return Object; }
  void Set(vtkInformationDataObjectKey*, vtkDataObject* o) { Object = o; }
  int* Get(vtkInformationIntegerVectorKey*) { return Whole; }
  void Set(vtkInformationIntegerVectorKey*, const int* v, int n) { for (int a = 0; a < n && a < 6; a++) Whole[a] = v[a]; }
private:
  vtkDataObject* Object;
  int Whole[6];
};
class vtkInformationVector : public vtkObject
{
public:
  static vtkInformationVector* New() { return new vtkInformationVector; }
  vtkInformationVector() : Info(vtkInformation::New()) {}
  ~vtkInformationVector() { Info->Delete(); }
  vtkInformation* GetInformationObject(int) { return Info; }
private:
  vtkInformation* Info;
};
class vtkStreamingDemandDrivenPipeline
{
public:
  static vtkInformationIntegerVectorKey* WHOLE_EXTENT() { static vtkInformationIntegerVectorKey k; return &k; }
};

/* vtkImageAlgorithm: one input port, one vtkImageData output; Update() = RequestInformation + RequestData */
class vtkImageAlgorithm : public vtkObject
{
public:
  vtkImageAlgorithm() : Input(0), Output(vtkImageData::New()), LastStatus(0) {}
  ~vtkImageAlgorithm() { Output->Delete(); if (Input) Input->UnRegister(this); }
  void SetNumberOfInputPorts(int) {}
  void SetInputData(vtkImageData* in) { if (in) in->Register(this); if (Input) Input->UnRegister(this); Input = in; }
  void SetInputData(int, vtkImageData* in) { SetInputData(in); }
  vtkImageData* GetOutput() { return Output; }
  virtual void Update()
  {
    vtkNew<vtkInformationVector> inV, outV;
    vtkInformationVector* ins[1] = {inV.Get()};
    inV->GetInformationObject(0)->Set(vtkDataObject::DATA_OBJECT(), Input);
    if (Input) inV->GetInformationObject(0)->Set(vtkStreamingDemandDrivenPipeline::WHOLE_EXTENT(), Input->GetExtent(), 6);
    outV->GetInformationObject(0)->Set(vtkDataObject::DATA_OBJECT(), Output);
    vtkInformation* req = 0;
    this->RequestInformation(req, ins, outV.Get());
    this->RequestUpdateExtent(req, ins, outV.Get());
    LastStatus = this->RequestData(req, ins, outV.Get());
  }
  int GetLastRequestDataStatus() const { return LastStatus; }     /* stand-in only */
protected:
  virtual int RequestData(vtkInformation*, vtkInformationVector**, vtkInformationVector*) { return 1; }
  virtual int RequestInformation(vtkInformation*, vtkInformationVector**, vtkInformationVector*) { return 1; }
  virtual int RequestUpdateExtent(vtkInformation*, vtkInformationVector**, vtkInformationVector*) { return 1; }
  vtkImageData* Input;
  vtkImageData* Output;
  int LastStatus;
};

/* ---- vtksys ------------------------------------------------------------------------------------------ */

namespace vtksys {
class SystemTools
{
public:
  static void ConvertToUnixSlashes(std::string& p) { for (size_t i = 0; i < p.size(); i++) if (p[i] == '\\') p[i] = '/'; }
  static std::string GetCurrentWorkingDirectory()
  {
    char buf[4096];
    return getcwd(buf, sizeof(buf)) ? std::string(buf) : std::string();
  }
};
}  // namespace vtksys

#endif /* DMI_VTK_STANDIN_H */
