// VTK-free readers for the reference's list files and .krtd camera files, so that the host classes in this
// directory consume the same on-disk inputs as the reference.  Behaviour follows Sources/Helper.h (cited per
// function); the code is written for this library (no VTK types, no element-wise matrix setters).  Header-only.
#pragma once
#include <array>
#include <fstream>
#include <iostream>
#include <sstream>
#include <string>
#include <vector>
#include <unistd.h>

namespace dmihost {
namespace help {

// Directory part of a path, '\\' treated as '/': "" when there is no separator, "/" for a file in the root,
// "C:/" for a drive root (the rules of Helper.h:32-55).
inline std::string GetFilenamePath(std::string path)
{
  for (size_t q = 0; q < path.size(); q++)
    if (path[q] == '\\') path[q] = '/';
  const size_t cut = path.rfind('/');
  if (cut == std::string::npos) return std::string();
  if (cut == 0) return "/";
  if (cut == 2 && path[1] == ':') return path.substr(0, 2) + "/";
  return path.substr(0, cut);
}

// The text after the last blank of a line, with the reference's tokenisation (Helper.h:18-27, 83-95): blanks
// separate tokens, one trailing blank does not open a new token, two do (an empty one).  false: empty line.
inline bool LastBlankSeparatedToken(const std::string& line, std::string& token)
{
  bool any = false;
  for (size_t pos = 0; pos < line.size();)
  {
    const size_t blank = line.find(' ', pos);
    any = true;
    if (blank == std::string::npos) { token.assign(line, pos, std::string::npos); break; }
    token.assign(line, pos, blank - pos);
    pos = blank + 1;
  }
  return any;
}

// A list file names one data file per line (its last blank-separated token), relative to the directory of the
// list file itself, or to the working directory when the list was given without one (Helper.h:60-100).
inline std::vector<std::string> ExtractAllFilePath(const char* listFile)
{
  std::vector<std::string> files;
  std::ifstream in(listFile);
  if (!in.is_open())
  {
    std::cerr << "Unable to open : " << listFile << std::endl;
    return files;
  }
  std::string dir = GetFilenamePath(listFile);
  if (dir.empty())
  {
    char cwd[4096];
    dir = getcwd(cwd, sizeof(cwd)) ? cwd : ".";
  }
  for (std::string line, name; std::getline(in, line);)
  {
    if (!line.empty() && line[line.size() - 1] == '\r') line.erase(line.size() - 1);      // CRLF lists
    if (LastBlankSeparatedToken(line, name)) files.push_back(dir + "/" + name);
  }
  return files;
}

// .krtd layout (Helper.h:105-168): K on lines 1-3, line 4 ignored, R on lines 5-7, line 8 ignored, T on line 9;
// three numbers are taken from each of those lines, a missing one reads as 0.  Outputs are the row-major 4x4
// matrices the C ABI takes: K4 = K inside an identity (ReconstructionData.cxx:199-209), RT4 = [R | T; 0 0 0 1].
inline bool ReadKrtdFile(const std::string& filename, double K4[16], double RT4[16])
{
  std::ifstream in(filename.c_str());
  if (!in.is_open())
  {
    std::cerr << "Unable to open krtd file : " << filename << std::endl;
    return false;
  }
  auto three = [&in]() {
    std::array<double, 3> v = {{0.0, 0.0, 0.0}};
    std::string line;
    std::getline(in, line);
    std::istringstream numbers(line);
    for (double& x : v)
      if (!(numbers >> x)) x = 0.0;
    return v;
  };
  auto skipLine = [&in]() { std::string ignored; std::getline(in, ignored); };

  std::array<double, 3> Krow[3], Rrow[3];
  for (auto& r : Krow) r = three();
  skipLine();
  for (auto& r : Rrow) r = three();
  skipLine();
  const std::array<double, 3> T = three();

  for (int r = 0; r < 4; r++)
    for (int c = 0; c < 4; c++)
    {
      K4[4 * r + c] = (r < 3 && c < 3) ? Krow[r][c] : (r == c ? 1.0 : 0.0);
      RT4[4 * r + c] = (r < 3) ? (c < 3 ? Rrow[r][c] : T[r]) : (c == 3 ? 1.0 : 0.0);
    }
  return true;
}

}  // namespace help
}  // namespace dmihost
