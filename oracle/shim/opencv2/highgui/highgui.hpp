// opencv2/highgui/highgui.hpp -- part of the OpenCV API shim of oracle/ (see ../opencv.hpp).  The reference includes it
// (Hypothesis.h:33) but uses nothing from it on the paths compiled into oracle/_ref.
#pragma once
#include "../opencv.hpp"
