// TEST INFRASTRUCTURE ONLY - what Thirdparty/DBoW2 needs from <opencv2/core/core.hpp> beyond the oracle's cv::Mat shim:
// the standard headers the real header pulls in, and cv::FileStorage / cv::FileNode declarations for the YAML save / load
// members of TemplatedVocabulary (virtual, so their bodies must compile; the reference loads its vocabulary with
// loadFromTextFile (src/System.cc:132) and nothing here ever opens a FileStorage: isOpened() is false).
#pragma once
#include <sstream>
#include <string>
#include <opencv2/cvshim.hpp>
#ifndef CV_32F
#define CV_32F 5   // only named by FORB::toMat32F, which nothing calls
#endif
namespace cv {
struct FileNode {
  FileNode operator[](const char*) const { return FileNode(); }
  FileNode operator[](const std::string&) const { return FileNode(); }
  FileNode operator[](int) const { return FileNode(); }
  size_t size() const { return 0; }
  operator int() const { return 0; }
  operator double() const { return 0; }
  operator std::string() const { return std::string(); }
};
struct FileStorage {
  enum { READ = 0, WRITE = 1 };
  FileStorage(const char*, int) {}
  bool isOpened() const { return false; }
  void release() {}
  FileNode operator[](const std::string&) const { return FileNode(); }
  FileNode operator[](const char*) const { return FileNode(); }
};
template <class T>
FileStorage& operator<<(FileStorage& f, const T&) { return f; }
}  // namespace cv
