// VTK-free restatement of the reference's file-list and .krtd parsing (Sources/Helper.h), so that the
// host classes in this directory read the same on-disk inputs as the reference.  Header-only.
#pragma once
#include <fstream>
#include <iostream>
#include <sstream>
#include <string>
#include <vector>
#include <unistd.h>

namespace dmihost {
namespace help {

// Helper.h:18-27
inline void SplitString(const std::string& s, char delim, std::vector<std::string>& elems)
{
  std::stringstream ss(s);
  std::string item;
  while (std::getline(ss, item, delim)) elems.push_back(item);
}

// Helper.h:32-55 (vtksys::SystemTools::ConvertToUnixSlashes reduced to the backslash replacement)
inline std::string GetFilenamePath(const std::string& filename)
{
  std::string fn = filename;
  for (char& ch : fn) if (ch == '\\') ch = '/';
  const std::string::size_type slash_pos = fn.rfind("/");
  if (slash_pos == std::string::npos) return "";
  std::string ret = fn.substr(0, slash_pos);
  if (ret.size() == 2 && ret[1] == ':') return ret + '/';
  if (ret.empty()) return "/";
  return ret;
}

// Helper.h:60-100: one path per line, the LAST space-separated token of the line, resolved against
// the directory of the list file (or the working directory).
inline std::vector<std::string> ExtractAllFilePath(const char* globalPath)
{
  std::vector<std::string> pathList;
  std::ifstream container(globalPath);
  if (!container.is_open())
  {
    std::cerr << "Unable to open : " << globalPath << std::endl;
    return pathList;
  }
  std::string directoryPath = GetFilenamePath(std::string(globalPath));
  if (directoryPath == "")
  {
    char buf[4096];
    directoryPath = getcwd(buf, sizeof(buf)) ? std::string(buf) : std::string(".");
  }
  std::string path;
  while (!container.eof())
  {
    std::getline(container, path);
    if (!path.empty() && path.back() == '\r') path.pop_back();
    std::vector<std::string> elems;
    SplitString(path, ' ', elems);
    if (elems.size() == 0) continue;
    pathList.push_back(directoryPath + "/" + elems[elems.size() - 1]);
  }
  return pathList;
}

// Helper.h:105-168: 3 lines K, 1 line skipped, 3 lines R, 1 line skipped, 1 line T.
// K4 = the 3x3 inside an identity 4x4 (ReconstructionData.cxx:199-209); RT4 = [R | T], last row 0 0 0 1.
inline bool ReadKrtdFile(const std::string& filename, double K4[16], double RT4[16])
{
  std::ifstream file(filename.c_str());
  if (!file.is_open())
  {
    std::cerr << "Unable to open krtd file : " << filename << std::endl;
    return false;
  }
  for (int i = 0; i < 16; i++) { K4[i] = (i % 5 == 0) ? 1.0 : 0.0; RT4[i] = 0.0; }
  std::string line;
  for (int i = 0; i < 3; i++)
  {
    getline(file, line);
    std::istringstream iss(line);
    for (int j = 0; j < 3; j++) { double value = 0; iss >> value; K4[4 * i + j] = value; }
  }
  getline(file, line);
  for (int i = 0; i < 3; i++)
  {
    getline(file, line);
    std::istringstream iss(line);
    for (int j = 0; j < 3; j++) { double value = 0; iss >> value; RT4[4 * i + j] = value; }
  }
  getline(file, line);
  getline(file, line);
  std::istringstream iss(line);
  for (int i = 0; i < 3; i++) { double value = 0; iss >> value; RT4[4 * i + 3] = value; }
  RT4[12] = RT4[13] = RT4[14] = 0.0;
  RT4[15] = 1.0;
  return true;
}

}  // namespace help
}  // namespace dmihost
