// cvshim_filestorage.h — compile-only stand-ins for cv::FileStorage / cv::FileNode so the reference's
// third_part/DBoW2/DBoW2/TemplatedVocabulary.h (virtual save/load members) instantiates without
// OpenCV.  ORACLE / TEST INFRASTRUCTURE ONLY.  Nothing here is ever executed: the oracle loads
// vocabularies through the reference's own loadFromTextFile (TemplatedVocabulary.h:1338-1420).
#ifndef TB_CVSHIM_FILESTORAGE_H
#define TB_CVSHIM_FILESTORAGE_H
#include <cstdlib>
#include <string>
namespace cv {
struct FileNode {
    FileNode operator[](const std::string&) const { std::abort(); }
    FileNode operator[](const char*) const { std::abort(); }
    FileNode operator[](int) const { std::abort(); }
    size_t size() const { std::abort(); }
    operator int() const { std::abort(); }
    operator double() const { std::abort(); }
    operator std::string() const { std::abort(); }
};
struct FileStorage {
    enum { READ = 0, WRITE = 1 };
    FileStorage(const char*, int) {}
    bool isOpened() const { return false; }
    FileNode operator[](const std::string&) const { std::abort(); }
};
template <typename T> inline FileStorage& operator<<(FileStorage& fs, const T&) { std::abort(); return fs; }
}  // namespace cv
#endif
