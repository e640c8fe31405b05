// Stand-in for rpg_vikit vision.h — nothing from it is used on the hot path.
#ifndef PLSVO_REFDEPS_VIKIT_VISION
#define PLSVO_REFDEPS_VIKIT_VISION
#include <opencv2/opencv.hpp>
#endif
