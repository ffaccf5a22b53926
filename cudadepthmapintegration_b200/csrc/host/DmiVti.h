// Minimal VTK-free reader / writer for the depth-map .vti files the reference loads with
// vtkXMLImageDataReader (Sources/ReconstructionData.cxx:223-229): point-data arrays "Depths",
// "Best Cost Values" (Float64 / Float32) and "Color" (UInt8 x 3), array names at
// ReconstructionData.cxx:95,144,146.  Supported encodings: format="ascii" and format="appended" with
// <AppendedData encoding="raw"> and no compressor (header_type UInt32 or UInt64).  Anything else
// (base64, zlib) is reported as an error -- convert such files with VTK, or link the VTK adapter.
// Header-only.
#pragma once
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <iostream>
#include <sstream>
#include <string>
#include <vector>

namespace dmihost {

struct DepthMapImage
{
  int W = 0, H = 0;
  std::vector<double> depths;     // H*W, bottom-up rows (VTK point order)
  std::vector<double> bestCost;   // H*W or empty
  std::vector<uint8_t> color;     // H*W*3 or empty
};

namespace vti_detail {

inline std::string attr(const std::string& tag, const std::string& name)
{
  const std::string key = name + "=\"";
  size_t p = 0;
  while ((p = tag.find(key, p)) != std::string::npos)
  {
    if (p == 0 || tag[p - 1] == ' ' || tag[p - 1] == '\t' || tag[p - 1] == '\n')
    {
      const size_t b = p + key.size(), e = tag.find('"', b);
      return e == std::string::npos ? "" : tag.substr(b, e - b);
    }
    p += key.size();
  }
  return "";
}

template <typename T>
inline void convert(const char* raw, size_t n, std::vector<double>& out)
{
  out.resize(n);
  for (size_t i = 0; i < n; i++) { T v; memcpy(&v, raw + i * sizeof(T), sizeof(T)); out[i] = (double)v; }
}

}  // namespace vti_detail

inline bool ReadVti(const std::string& path, DepthMapImage& img, std::string& err)
{
  using namespace vti_detail;
  std::ifstream f(path.c_str(), std::ios::binary);
  if (!f.is_open()) { err = "Unable to open : " + path; return false; }
  std::string s((std::istreambuf_iterator<char>(f)), std::istreambuf_iterator<char>());
  const size_t vf = s.find("<VTKFile");
  if (vf == std::string::npos) { err = path + ": not a VTK XML file"; return false; }
  const std::string vtag = s.substr(vf, s.find('>', vf) - vf);
  if (attr(vtag, "type") != "ImageData") { err = path + ": VTKFile type is not ImageData"; return false; }
  if (!attr(vtag, "compressor").empty()) { err = path + ": compressed .vti is not supported by the VTK-free reader"; return false; }
  if (!attr(vtag, "byte_order").empty() && attr(vtag, "byte_order") != "LittleEndian") { err = path + ": big-endian .vti not supported"; return false; }
  const bool hdr64 = attr(vtag, "header_type") == "UInt64";
  const size_t id = s.find("<ImageData");
  if (id == std::string::npos) { err = path + ": no <ImageData>"; return false; }
  const std::string itag = s.substr(id, s.find('>', id) - id);
  int e[6] = {0, -1, 0, -1, 0, 0};
  { std::istringstream iss(attr(itag, "WholeExtent")); for (int i = 0; i < 6; i++) iss >> e[i]; }
  img.W = e[1] - e[0] + 1; img.H = e[3] - e[2] + 1;
  if (img.W <= 0 || img.H <= 0 || e[5] != e[4]) { err = path + ": expected a 2-D image extent"; return false; }
  const size_t npix = (size_t)img.W * img.H;
  // appended data block
  const char* app = nullptr;
  const size_t ad = s.find("<AppendedData");
  if (ad != std::string::npos)
  {
    const std::string atag = s.substr(ad, s.find('>', ad) - ad);
    const size_t us = s.find('_', s.find('>', ad));
    if (us != std::string::npos) app = s.data() + us + 1;
    if (attr(atag, "encoding") != "raw") app = nullptr;
  }
  const size_t pd0 = s.find("<PointData"), pd1 = s.find("</PointData>");
  if (pd0 == std::string::npos || pd1 == std::string::npos) { err = path + ": no <PointData>"; return false; }
  size_t p = pd0;
  img.depths.clear(); img.bestCost.clear(); img.color.clear();
  while ((p = s.find("<DataArray", p)) != std::string::npos && p < pd1)
  {
    const size_t te = s.find('>', p);
    const std::string tag = s.substr(p, te - p);
    const std::string name = attr(tag, "Name"), type = attr(tag, "type"), format = attr(tag, "format");
    int comps = attr(tag, "NumberOfComponents").empty() ? 1 : atoi(attr(tag, "NumberOfComponents").c_str());
    const bool isColor = name == "Color";
    if (name == "Depths" || name == "Best Cost Values" || isColor)
    {
      const size_t count = npix * (size_t)comps;
      std::vector<double> vals;
      std::vector<uint8_t> bytes;
      if (format == "ascii")
      {
        const size_t de = s.find("</DataArray>", te);
        std::istringstream iss(s.substr(te + 1, de - te - 1));
        if (isColor) { bytes.resize(count); for (size_t i = 0; i < count; i++) { int v = 0; iss >> v; bytes[i] = (uint8_t)v; } }
        else { vals.resize(count); for (size_t i = 0; i < count; i++) iss >> vals[i]; }
      }
      else if (format == "appended")
      {
        if (!app) { err = path + ": appended data must be encoding=\"raw\" for the VTK-free reader"; return false; }
        const size_t off = (size_t)strtoull(attr(tag, "offset").c_str(), nullptr, 10);
        const char* q = app + off;
        uint64_t nbytes = 0;
        if (hdr64) { memcpy(&nbytes, q, 8); q += 8; } else { uint32_t n32; memcpy(&n32, q, 4); nbytes = n32; q += 4; }
        const size_t esz = type == "Float64" ? 8 : type == "Float32" ? 4 : type == "UInt8" ? 1 : 0;
        if (esz == 0 || nbytes != count * esz || q + nbytes > s.data() + s.size())
        { err = path + ": array '" + name + "' has an unsupported type or size"; return false; }
        if (isColor) { bytes.assign((const uint8_t*)q, (const uint8_t*)q + nbytes); }
        else if (type == "Float64") convert<double>(q, count, vals);
        else convert<float>(q, count, vals);
      }
      else { err = path + ": format '" + format + "' is not supported by the VTK-free reader"; return false; }
      if (name == "Depths") img.depths.swap(vals);
      else if (name == "Best Cost Values") img.bestCost.swap(vals);
      else { if (comps != 3 || type != "UInt8") { err = path + ": Color must be UInt8 x 3"; return false; } img.color.swap(bytes); }
    }
    p = te;
  }
  if (img.depths.size() != npix) { err = path + ": no 'Depths' array"; return false; }
  return true;
}

inline bool WriteVti(const std::string& path, const DepthMapImage& img)
{
  std::ofstream f(path.c_str(), std::ios::binary);
  if (!f.is_open()) return false;
  const size_t npix = (size_t)img.W * img.H;
  size_t off = 0;
  f << "<?xml version=\"1.0\"?>\n<VTKFile type=\"ImageData\" version=\"0.1\" byte_order=\"LittleEndian\" header_type=\"UInt64\">\n";
  f << "  <ImageData WholeExtent=\"0 " << img.W - 1 << " 0 " << img.H - 1 << " 0 0\" Origin=\"0 0 0\" Spacing=\"1 1 1\">\n";
  f << "    <Piece Extent=\"0 " << img.W - 1 << " 0 " << img.H - 1 << " 0 0\">\n      <PointData>\n";
  f << "        <DataArray type=\"Float64\" Name=\"Depths\" format=\"appended\" offset=\"" << off << "\"/>\n";
  off += 8 + npix * 8;
  if (img.bestCost.size() == npix)
  {
    f << "        <DataArray type=\"Float64\" Name=\"Best Cost Values\" format=\"appended\" offset=\"" << off << "\"/>\n";
    off += 8 + npix * 8;
  }
  if (img.color.size() == npix * 3)
    f << "        <DataArray type=\"UInt8\" Name=\"Color\" NumberOfComponents=\"3\" format=\"appended\" offset=\"" << off << "\"/>\n";
  f << "      </PointData>\n    </Piece>\n  </ImageData>\n  <AppendedData encoding=\"raw\">\n   _";
  auto block = [&](const void* data, uint64_t nbytes) { f.write((const char*)&nbytes, 8); f.write((const char*)data, (std::streamsize)nbytes); };
  block(img.depths.data(), npix * 8);
  if (img.bestCost.size() == npix) block(img.bestCost.data(), npix * 8);
  if (img.color.size() == npix * 3) block(img.color.data(), npix * 3);
  f << "\n  </AppendedData>\n</VTKFile>\n";
  return (bool)f;
}

}  // namespace dmihost
