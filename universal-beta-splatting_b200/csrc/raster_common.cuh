// Constants shared by the compositing kernels.
#pragma once
#include <stdint.h>

namespace ubs {
constexpr int kTile = 16;                   // tile edge in pixels (the reference caller always uses 16)
constexpr int kTilePixels = kTile * kTile;  // one thread per pixel, one CTA per tile
}  // namespace ubs
