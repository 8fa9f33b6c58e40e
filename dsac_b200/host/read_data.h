// read_data.h -- the reference's ground-truth reader for 7-Scenes pose files (core/read_data.h:74, core/read_data.cpp:69-133),
// without OpenCV / png++.  Only the pose-file convention is mirrored (the frames themselves never enter the hypothesis
// engine: its input is the scene-coordinate grid).
#pragma once
#include <string>

#include "types.h"

namespace jp {
// Reads a 7-Scenes "frame-XXXXXX.pose.txt" (4x4 camera-to-world, metres; only the first three rows are used), subtracts
// the scene offset of ./translation.txt if that file exists (warns otherwise), flips the y and z axes, inverts, and
// fills info.rotation / info.center (scene -> camera, metres) -- in float, like the reference.  Returns false (and
// info.visible = false) if the file cannot be opened.
bool readData(const std::string infoFile, jp::info_t& info);
}  // namespace jp
