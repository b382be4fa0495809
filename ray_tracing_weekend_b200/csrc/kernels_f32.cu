// kernels_f32.cu — the fast (FP32) instantiation of every kernel.  Compiled with FMA contraction on.
#include <algorithm>
#include "rtw_launch.cuh"
namespace rtw { RTW_DEFINE_LAUNCHERS(f32, float, false) }
