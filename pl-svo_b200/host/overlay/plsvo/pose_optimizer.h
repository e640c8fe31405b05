// Drop-in replacement of include/plsvo/pose_optimizer.h: both plsvo::pose_optimizer::optimizeGaussNewton overloads come
// from the B200 shim.
#pragma once
#include <plsvo_shim.h>
