// Track ingestion on the host: waypoint file -> arc-length re-parametrised 6-D spline table.
// Mirrors Track (cpp/src/Params/track.cpp) and ArcLengthSpline::gen6DSpline / fitSpline
// (cpp/src/Spline/arc_length_spline.cpp:213-265) of the reference; the device consumes the
// resulting TrackTable (mpcc_types.h).
#pragma once
#include "../mpcc_types.h"
#include <string>
#include <vector>

namespace mpcc {

struct Waypoints {
    std::vector<double> X, Y, Z;
    std::vector<double> R;  // n x 9, row-major rotation matrices
    size_t size() const { return X.size(); }
};

// Track::Track (track.cpp:19-54): X/Y/Z + quaternion arrays -> rotation matrices
Waypoints load_track_json(const std::string& file);
// Track::getTrack (track.cpp:56-66): translate so that the first waypoint is init_position
void shift_track(Waypoints& w, const double init_position[3]);

// fitSpline: chord-length fit -> resample 100 -> re-measure -> refit -> resample 100 -> regular spline
void fit_track(const Waypoints& w, TrackTable& out);

}  // namespace mpcc
