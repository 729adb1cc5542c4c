// Build-harness stub (oracle/_ref only): the reference includes the third-party header
// libnpy (llohse/libnpy, pre-1.0 API; not vendored upstream, no version pinned).  Only the
// declarations are needed: the harness never reaches the reference's file I/O.
#pragma once
#include <string>
#include <vector>
namespace npy {
typedef unsigned long ndarray_len_t;
template <class S> inline void LoadArrayFromNumpy(const std::string&, std::vector<ndarray_len_t>&, std::vector<S>&) {}
template <class S> inline void SaveArrayAsNumpy(const std::string&, bool, unsigned, const unsigned long*, const S*) {}
template <class S> inline void SaveArrayAsNumpy(const std::string&, bool, unsigned, const unsigned long*, const std::vector<S>&) {}
}
