/* Minimal stand-in for CImg.h so that the UNMODIFIED reference image3d.cpp compiles
 * without libpng/X11.  Only the members image3d.cpp:103-127 touches are declared;
 * the oracle driver never calls image3d::load (it fills the public _dim/_voxels
 * fields directly), so load() just fails loudly.  Test infrastructure only. */
#ifndef ORACLE_STUB_CIMG_H
#define ORACLE_STUB_CIMG_H
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cmath>
#include <cassert>
#include <sstream>
#include <exception>
namespace cimg_library {
template <typename T> struct CImg {
    CImg &load(const char *path) {
        std::fprintf(stderr, "oracle CImg stub: load(%s) is not available\n", path);
        std::abort();
        return *this;
    }
    int width() const { return 0; }
    int height() const { return 0; }
    CImg &noise(double, int) { return *this; }
    T operator()(int, int) const { return T(0); }
};
}
#endif
