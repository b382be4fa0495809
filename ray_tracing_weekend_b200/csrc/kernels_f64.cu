// kernels_f64.cu — the reference-exact (f64) instantiation.  MUST be compiled with -fmad=false: the
// reference (Rust) never contracts a*b+c into an FMA, and bit-compatibility depends on it.
#include <algorithm>
#include "rtw_launch.cuh"
namespace rtw { RTW_DEFINE_LAUNCHERS(f64, double, true) }
