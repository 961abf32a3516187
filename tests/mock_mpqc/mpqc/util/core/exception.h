// MOCK of util/core/exception.h constructors used by the adapter (signatures as in the reference, :135-612)
#pragma once
#include <stdexcept>
namespace mpqc {
struct Exception : std::runtime_error { explicit Exception(const char* d) : std::runtime_error(d ? d : "") {} };
struct ProgrammingError : Exception { ProgrammingError(const char* d = 0, const char* = 0, int = 0) : Exception(d) {} };
struct InputError : Exception { InputError(const char* d = 0, const char* = 0, int = 0, const char* = 0, const char* = 0) : Exception(d) {} };
struct MemAllocFailed : Exception { MemAllocFailed(const char* d = 0, const char* = 0, int = 0, size_t = 0) : Exception(d) {} };
struct FeatureDisabled : Exception { FeatureDisabled(const char* d = 0, const char* = 0, int = 0) : Exception(d) {} };
struct FileOperationFailed : Exception {
  enum FileOperation { Unknown, OpenR, OpenW, OpenRW };
  FileOperationFailed(const char* d = 0, const char* = 0, int = 0, const char* = 0, FileOperation = Unknown) : Exception(d) {}
};
}  // namespace mpqc
