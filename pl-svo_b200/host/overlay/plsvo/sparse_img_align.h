// Drop-in replacement of include/plsvo/sparse_img_align.h: put this directory in front of the reference's include path
//  and plsvo::SparseImgAlign comes from the B200 shim; src/frame_handler_mono.cpp compiles unchanged.
#pragma once
#include <plsvo_shim.h>
