// TEST INFRASTRUCTURE (oracle): extern "C" door onto the reference's own CTW implementation, compiled from the sources
// where they lie under /root/reference/chaos (cppctw.cpp, cppctw.h) by oracle/Makefile into oracle/_ref/libctw_ref.so.
// Nothing of the reference is copied: this file only calls estimate_entropy() declared in the reference's cppctw.h.
#include <vector>

#include "cppctw.h"

extern "C" double ctw_ref_estimate_entropy(const signed char* seq, long long n, int alphabet_size) {
  std::vector<char> v(seq, seq + n);
  return estimate_entropy(v, (char)alphabet_size);
}
