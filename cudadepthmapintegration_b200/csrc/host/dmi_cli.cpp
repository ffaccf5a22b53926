// dmi_cli: VTK-free command line over the host classes, with the reference CLIs' flag names.
//
//   dmi_cli reconstruction --gridDims nx ny nz | --gridSpacing sx sy sz  --gridOrigin x y z --gridEnd x y z
//                          [--gridVecX ..] [--gridVecY ..] [--gridVecZ ..] --dataFolder DIR
//                          [--depthMapFile vtiList.txt] [--KRTFile kList.txt] --rayThick T --rayRho R
//                          --rayEta E --rayDelta D [--threshBestCost 0.14] [--forceCubicVoxel] [--verbose]
//                          --outputGridFilename out.mhd
//   (flags, defaults and checks of Reconstruction/main.cxx:63-83, 216-343; the grid matrix is built like
//    CreateGridMatrixFromInput :345-359.  The reference's isocontour / .vts / .vtp writers are VTK and stay
//    in the reference; this tool writes the cell scalars as MetaImage .mhd + .raw, cf. main.cxx:151-161.)
//
//   dmi_cli coloration --input points.f32 --output prefix --krtd kList.txt --vti vtiList.txt [--verbose]
//   (flags of Coloration/main.cxx:112-117; the mesh is its float32 xyz array; writes prefix.mean.u8,
//    prefix.median.u8, prefix.nb.i32 = MeanColoration, MedianColoration, NbProjectedDepthMap.)
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <map>

#include "DmiHostClasses.h"

using namespace dmihost;

typedef std::map<std::string, std::vector<std::string>> Args;

static Args parse(int argc, char** argv, int first)
{
  Args a;
  std::string cur;
  for (int i = first; i < argc; i++)
  {
    const std::string t = argv[i];
    const bool flag = t.size() > 2 && t[0] == '-' && t[1] == '-' && !(t[2] >= '0' && t[2] <= '9') && t[2] != '.';
    if (flag) { cur = t; a[cur]; }
    else if (!cur.empty()) a[cur].push_back(t);
  }
  return a;
}

static bool num3(const Args& a, const char* k, std::vector<double>& out)
{
  out.clear();
  auto it = a.find(k);
  if (it == a.end()) return false;
  for (const auto& s : it->second) out.push_back(atof(s.c_str()));
  return !out.empty();
}

static double num(const Args& a, const char* k, double def)
{
  auto it = a.find(k);
  return (it == a.end() || it->second.empty()) ? def : atof(it->second[0].c_str());
}

static std::string str(const Args& a, const char* k, const std::string& def)
{
  auto it = a.find(k);
  return (it == a.end() || it->second.empty()) ? def : it->second[0];
}

// --gpus N (new; the reference is single-GPU): devices 0 .. N-1
static std::vector<int> devices(const Args& a)
{
  std::vector<int> d;
  const int n = (int)num(a, "--gpus", 1);
  for (int q = 0; q < n; q++) d.push_back(q);
  return d;
}

static int reconstruction(const Args& a)
{
  std::vector<double> dimsv, spacing, origin, end, vx, vy, vz;
  num3(a, "--gridDims", dimsv); num3(a, "--gridSpacing", spacing);
  if (!num3(a, "--gridOrigin", origin) || origin.size() != 3 || !num3(a, "--gridEnd", end) || end.size() != 3)
  { std::cerr << "Error arguments." << std::endl; return 1; }
  if (!spacing.empty() && !dimsv.empty()) { std::cerr << "Error : Spacing and dimensions can't be both set" << std::endl; return 1; }
  if (dimsv.size() == 1) { dimsv.push_back(dimsv[0]); dimsv.push_back(dimsv[0]); }
  const double thick = num(a, "--rayThick", 2), rho = num(a, "--rayRho", 0.8), eta = num(a, "--rayEta", 0.03);
  const double delta = num(a, "--rayDelta", 0.3), thresh = num(a, "--threshBestCost", 0.14);
  const std::string out = str(a, "--outputGridFilename", ""), folder = str(a, "--dataFolder", ".");
  const std::string vti = str(a, "--depthMapFile", "vtiList.txt"), krt = str(a, "--KRTFile", "kList.txt");
  if (out.empty() || delta < thick || eta < 0 || eta > 1) { std::cerr << "Error arguments." << std::endl; return 1; }   // main.cxx:270-271
  if (!num3(a, "--gridVecX", vx)) vx = {1, 0, 0};
  if (!num3(a, "--gridVecY", vy)) vy = {0, 1, 0};
  if (!num3(a, "--gridVecZ", vz)) vz = {0, 0, 1};
  auto dot = [](const std::vector<double>& p, const std::vector<double>& q) { return p[0] * q[0] + p[1] * q[1] + p[2] * q[2]; };
  if (std::fabs(dot(vx, vy)) > 10e-6 || std::fabs(dot(vy, vz)) > 10e-6 || std::fabs(dot(vz, vx)) > 10e-6)
  { std::cerr << "Given vectors are not orthogonals." << std::endl; return 1; }                                          // :363-382
  const double size[3] = {end[0] - origin[0], end[1] - origin[1], end[2] - origin[2]};
  if (spacing.empty()) { if (dimsv.size() != 3) { std::cerr << "Error arguments." << std::endl; return 1; } spacing = {size[0] / dimsv[0], size[1] / dimsv[1], size[2] / dimsv[2]}; }
  if (dimsv.empty()) dimsv = {(double)(int)(size[0] / spacing[0]), (double)(int)(size[1] / spacing[1]), (double)(int)(size[2] / spacing[2])};
  if (a.count("--forceCubicVoxel")) { const double m = std::min(spacing[0], std::min(spacing[1], spacing[2])); spacing = {m, m, m}; }
  const int dims[3] = {(int)dimsv[0], (int)dimsv[1], (int)dimsv[2]};          // POINT dims of the vtkImageData, main.cxx:123
  double gm[16] = {vx[0], vx[1], vx[2], 0, vy[0], vy[1], vy[2], 0, vz[0], vz[1], vz[2], 0, 0, 0, 0, 1};   // :345-359

  CudaReconstructionFilter f;
  if (!f.SetDevices(devices(a))) return 1;
  f.SetInputGrid(dims, origin.data(), spacing.data());
  f.SetGridMatrix(gm);
  f.SetFilePathKRTD(folder + "/" + krt);
  f.SetFilePathVTI(folder + "/" + vti);
  f.SetRayPotentialThickness(thick); f.SetRayPotentialRho(rho); f.SetRayPotentialEta(eta); f.SetRayPotentialDelta(delta);
  f.SetThresholdBestCost(thresh);
  if (!f.Update() || f.GetOutput().empty()) return 1;
  if (a.count("--verbose")) std::cout << "Reconstruction time : " << f.GetExecutionTime() << " s" << std::endl;
  std::string raw = out;
  const size_t dotp = raw.rfind('.');
  raw = (dotp == std::string::npos ? raw : raw.substr(0, dotp)) + ".raw";
  {
    std::ofstream h(out.c_str());
    const std::string rawName = raw.substr(raw.rfind('/') == std::string::npos ? 0 : raw.rfind('/') + 1);
    h << "ObjectType = Image\nNDims = 3\nBinaryData = True\nBinaryDataByteOrderMSB = False\n"
      << "DimSize = " << dims[0] - 1 << " " << dims[1] - 1 << " " << dims[2] - 1 << "\n"
      << "ElementSpacing = " << spacing[0] << " " << spacing[1] << " " << spacing[2] << "\n"
      << "Offset = " << origin[0] + 0.5 * spacing[0] << " " << origin[1] + 0.5 * spacing[1] << " " << origin[2] + 0.5 * spacing[2] << "\n"
      << "ElementType = MET_DOUBLE\nElementDataFile = " << rawName << "\n";
    std::ofstream r(raw.c_str(), std::ios::binary);
    r.write((const char*)f.GetOutput().data(), (std::streamsize)(f.GetOutput().size() * 8));
    if (!h || !r) { std::cerr << "Unable to write " << out << std::endl; return 1; }
  }
  return 0;
}

static int coloration(const Args& a)
{
  const std::string in = str(a, "--input", ""), out = str(a, "--output", ""), krtd = str(a, "--krtd", ""), vti = str(a, "--vti", "");
  if (in.empty() || out.empty() || krtd.empty() || vti.empty()) { std::cerr << "Missing arguments..." << std::endl; return 1; }   // Coloration/main.cxx:127-132
  std::ifstream f(in.c_str(), std::ios::binary);
  if (!f.is_open()) { std::cerr << "Unable to open : " << in << std::endl; return 1; }
  std::vector<char> bytes((std::istreambuf_iterator<char>(f)), std::istreambuf_iterator<char>());
  std::vector<float> xyz(bytes.size() / 12 * 3);
  memcpy(xyz.data(), bytes.data(), xyz.size() * 4);
  MeshColoration mc(xyz, vti, krtd);
  if (!mc.SetDevices(devices(a))) return 1;
  if (!mc.ProcessColoration()) return 1;
  auto dump = [&](const std::string& suffix, const void* p, size_t n) { std::ofstream o((out + suffix).c_str(), std::ios::binary); o.write((const char*)p, (std::streamsize)n); return (bool)o; };
  if (!dump(".mean.u8", mc.MeanColoration.data(), mc.MeanColoration.size()) || !dump(".median.u8", mc.MedianColoration.data(), mc.MedianColoration.size()) ||
      !dump(".nb.i32", mc.NbProjectedDepthMap.data(), mc.NbProjectedDepthMap.size() * 4))
  { std::cerr << "Unable to write " << out << std::endl; return 1; }
  return 0;
}

// dmi_cli inspect --vti vtiList.txt --krtd kList.txt : parses the inputs (no GPU) and prints what it read
static int inspect(const Args& a)
{
  const std::vector<std::string> vti = help::ExtractAllFilePath(str(a, "--vti", "").c_str());
  const std::vector<std::string> krtd = help::ExtractAllFilePath(str(a, "--krtd", "").c_str());
  printf("views %zu krtd %zu\n", vti.size(), krtd.size());
  for (size_t v = 0; v < vti.size() && v < krtd.size(); v++)
  {
    DepthMapImage img;
    std::string err;
    double K[16], RT[16];
    if (!ReadVti(vti[v], img, err)) { std::cerr << err << std::endl; return 1; }
    if (!help::ReadKrtdFile(krtd[v], K, RT)) return 1;
    double sd = 0, sc = 0; unsigned long long scol = 0; size_t invalid = 0;
    for (double d : img.depths) { sd += d; invalid += d == -1.0; }
    for (double c : img.bestCost) sc += c;
    for (uint8_t c : img.color) scol += c;
    printf("view %zu W %d H %d depthsum %.17g costsum %.17g colorsum %llu invalid %zu K", v, img.W, img.H, sd, sc, scol, invalid);
    for (int i = 0; i < 16; i++) printf(" %.17g", K[i]);
    printf(" RT");
    for (int i = 0; i < 16; i++) printf(" %.17g", RT[i]);
    printf("\n");
  }
  return 0;
}

int main(int argc, char** argv)
{
  if (argc < 2) { std::cerr << "usage: dmi_cli reconstruction|coloration [flags]  (see the header of dmi_cli.cpp)" << std::endl; return 1; }
  const Args a = parse(argc, argv, 2);
  if (!strcmp(argv[1], "reconstruction")) return reconstruction(a);
  if (!strcmp(argv[1], "coloration")) return coloration(a);
  if (!strcmp(argv[1], "inspect")) return inspect(a);
  std::cerr << "unknown command " << argv[1] << std::endl;
  return 1;
}
