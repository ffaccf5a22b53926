// Minimal VTK-free reader / writer for the depth-map .vti files the reference loads with
// vtkXMLImageDataReader (Sources/ReconstructionData.cxx:223-229): point-data arrays "Depths",
// "Best Cost Values" (Float64 / Float32) and "Color" (UInt8 x 3), array names at
// ReconstructionData.cxx:95,144,146.  Supported: format="ascii"; format="binary" (inline base64);
// format="appended" with <AppendedData encoding="raw"> or encoding="base64"; header_type UInt32 / UInt64;
// compressor="vtkZLibDataCompressor" (the XML writers' defaults are appended + base64 + zlib).  Little-endian
// only.  Every array's decoded size is checked against the image extent.  The binary layouts follow the VTK XML
// format description (block header = byte count, or for compressed data: #blocks, block size, last block size,
// compressed sizes; headers base64-encoded on their own); VTK is not installed in this image, so files written
// by VTK itself have not been read here -- an unexpected layout is reported as an error, never guessed.
// Header-only; link zlib (-lz).
#pragma once
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <iostream>
#include <sstream>
#include <string>
#include <vector>
#include <zlib.h>

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

// base64 -> bytes; stops at the first character outside the alphabet ('=' padding, '<', blanks); returns the number
// of characters consumed through `used`
inline std::vector<unsigned char> Base64Decode(const char* p, size_t maxChars, size_t* used = nullptr)
{
  static signed char tab[256];
  static bool init = false;
  if (!init)
  {
    for (int i = 0; i < 256; i++) tab[i] = -1;
    const char* abc = "ABCDEFGHIJKLMNOPQRSTUVWXYZabcdefghijklmnopqrstuvwxyz0123456789+/";
    for (int i = 0; i < 64; i++) tab[(unsigned char)abc[i]] = (signed char)i;
    init = true;
  }
  std::vector<unsigned char> out;
  out.reserve(maxChars / 4 * 3 + 3);
  unsigned acc = 0;
  int bits = 0;
  size_t q = 0;
  for (; q < maxChars; q++)
  {
    const int v = tab[(unsigned char)p[q]];
    if (v < 0) break;
    acc = (acc << 6) | (unsigned)v;
    bits += 6;
    if (bits >= 8) { bits -= 8; out.push_back((unsigned char)((acc >> bits) & 0xff)); }
  }
  if (used) *used = q;
  return out;
}

inline size_t Base64Chars(size_t bytes) { return (bytes + 2) / 3 * 4; }

inline uint64_t HeaderWord(const unsigned char* p, bool hdr64)
{
  uint64_t v = 0;
  if (hdr64) memcpy(&v, p, 8); else { uint32_t w; memcpy(&w, p, 4); v = w; }
  return v;
}

// One binary data block (inline or appended) -> `want` bytes.  `p` points at the block, `avail` bounds it.
inline bool DecodeBlock(const char* p, size_t avail, bool base64, bool zlibCompressed, bool hdr64, size_t want,
                        std::vector<unsigned char>& out, std::string& why)
{
  const size_t hs = hdr64 ? 8 : 4;
  out.clear();
  if (!zlibCompressed)
  {
    if (!base64)
    {
      if (avail < hs) { why = "truncated block"; return false; }
      const uint64_t n = HeaderWord((const unsigned char*)p, hdr64);
      if (n != want || avail < hs + n) { why = "byte count does not match the image extent"; return false; }
      out.assign((const unsigned char*)p + hs, (const unsigned char*)p + hs + n);
      return true;
    }
    // the byte-count header is encoded on its own (padded) or together with the data: accept both
    size_t used = 0;
    const size_t hc = Base64Chars(hs);
    std::vector<unsigned char> h = Base64Decode(p, std::min(avail, hc), &used);
    if (h.size() < hs) { why = "truncated block header"; return false; }
    if (HeaderWord(h.data(), hdr64) != want) { why = "byte count does not match the image extent"; return false; }
    const bool separate = used < hc;                        // stopped at '=' padding
    if (separate)
    {
      if (avail < hc) { why = "truncated block"; return false; }
      out = Base64Decode(p + hc, std::min(avail - hc, Base64Chars(want)));
    }
    else
    {
      std::vector<unsigned char> all = Base64Decode(p, std::min(avail, Base64Chars(hs + want)));
      if (all.size() < hs + want) { why = "truncated block"; return false; }
      out.assign(all.begin() + (long)hs, all.begin() + (long)(hs + want));
    }
    if (out.size() < want) { why = "truncated block"; return false; }
    out.resize(want);
    return true;
  }
  // compressed: header = nBlocks, blockSize, lastBlockSize, compressedSize[nBlocks]
  std::vector<unsigned char> head;
  const unsigned char* data = nullptr;
  size_t dataAvail = 0;
  std::vector<unsigned char> decoded;
  uint64_t nb = 0;
  if (base64)
  {
    std::vector<unsigned char> first = Base64Decode(p, std::min(avail, Base64Chars(3 * hs)));
    if (first.size() < hs) { why = "truncated compression header"; return false; }
    nb = HeaderWord(first.data(), hdr64);
    if (nb > (1u << 24)) { why = "implausible number of compressed blocks"; return false; }
    const size_t hbytes = (3 + (size_t)nb) * hs, hchars = Base64Chars(hbytes);
    if (avail < hchars) { why = "truncated compression header"; return false; }
    head = Base64Decode(p, hchars);
    if (head.size() < hbytes) { why = "truncated compression header"; return false; }
    uint64_t total = 0;
    for (uint64_t b = 0; b < nb; b++) total += HeaderWord(head.data() + (3 + b) * hs, hdr64);
    decoded = Base64Decode(p + hchars, std::min(avail - hchars, Base64Chars((size_t)total)));
    if (decoded.size() < total) { why = "truncated compressed data"; return false; }
    data = decoded.data(); dataAvail = decoded.size();
  }
  else
  {
    if (avail < 3 * hs) { why = "truncated compression header"; return false; }
    nb = HeaderWord((const unsigned char*)p, hdr64);
    if (nb > (1u << 24) || avail < (3 + nb) * hs) { why = "truncated compression header"; return false; }
    head.assign((const unsigned char*)p, (const unsigned char*)p + (3 + nb) * hs);
    data = (const unsigned char*)p + (3 + nb) * hs; dataAvail = avail - (3 + nb) * hs;
  }
  const uint64_t blockSize = HeaderWord(head.data() + hs, hdr64), lastSize = HeaderWord(head.data() + 2 * hs, hdr64);
  const uint64_t expect = nb == 0 ? 0 : (nb - 1) * blockSize + (lastSize ? lastSize : blockSize);
  if (expect != want) { why = "uncompressed size does not match the image extent"; return false; }
  out.resize(want);
  size_t in = 0, off = 0;
  for (uint64_t b = 0; b < nb; b++)
  {
    const uint64_t cs = HeaderWord(head.data() + (3 + b) * hs, hdr64);
    const uint64_t us = (b + 1 == nb && lastSize) ? lastSize : blockSize;
    // sizes come from the file: written so that no sum of them can wrap
    if (cs > dataAvail - in || us > want - off) { why = "corrupt compression header"; return false; }
    uLongf got = (uLongf)us;
    if (uncompress(out.data() + off, &got, data + in, (uLong)cs) != Z_OK || got != us) { why = "zlib could not inflate a block"; return false; }
    in += cs; off += us;
  }
  return true;
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
  const std::string compressor = attr(vtag, "compressor");
  if (!compressor.empty() && compressor != "vtkZLibDataCompressor")
  { err = path + ": compressor " + compressor + " is not supported by the VTK-free reader (zlib only)"; return false; }
  const bool zcomp = !compressor.empty();
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
  bool appBase64 = false;
  const size_t ad = s.find("<AppendedData");
  if (ad != std::string::npos)
  {
    const std::string atag = s.substr(ad, s.find('>', ad) - ad);
    const size_t us = s.find('_', s.find('>', ad));
    if (us != std::string::npos) app = s.data() + us + 1;
    const std::string enc = attr(atag, "encoding");
    if (enc == "base64" || enc.empty()) appBase64 = true;      // base64 is the format's default
    else if (enc != "raw") app = nullptr;
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
        if (isColor) { bytes.resize(count); for (size_t i = 0; i < count && iss; i++) { int v = 0; iss >> v; bytes[i] = (uint8_t)v; } }
        else { vals.resize(count); for (size_t i = 0; i < count && iss; i++) iss >> vals[i]; }
        if (iss.fail()) { err = path + ": ascii array '" + name + "' is truncated or malformed"; return false; }
      }
      else if (format == "appended" || format == "binary")
      {
        const size_t esz = type == "Float64" ? 8 : type == "Float32" ? 4 : type == "UInt8" ? 1 : 0;
        if (esz == 0) { err = path + ": array '" + name + "' has an unsupported type"; return false; }
        const char* q = nullptr;
        bool b64 = true;
        if (format == "appended")
        {
          if (!app) { err = path + ": appended data section missing or in an unknown encoding"; return false; }
          const size_t off = (size_t)strtoull(attr(tag, "offset").c_str(), nullptr, 10);
          if (off >= (size_t)(s.data() + s.size() - app)) { err = path + ": array '" + name + "' points outside the file"; return false; }
          q = app + off;
          b64 = appBase64;
        }
        else
        {
          q = s.data() + te + 1;
          while (q < s.data() + s.size() && (*q == ' ' || *q == '\n' || *q == '\r' || *q == '\t')) q++;
        }
        std::vector<unsigned char> rawBytes;
        std::string why;
        if (!DecodeBlock(q, (size_t)(s.data() + s.size() - q), b64, zcomp, hdr64, count * esz, rawBytes, why))
        { err = path + ": array '" + name + "': " + why; return false; }
        if (isColor) bytes.assign(rawBytes.begin(), rawBytes.end());
        else if (type == "Float64") convert<double>((const char*)rawBytes.data(), count, vals);
        else convert<float>((const char*)rawBytes.data(), count, vals);
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
