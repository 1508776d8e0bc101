#pragma once
#define OPEN3D_VERSION "0.3.0"
namespace open3d { void PrintOpen3DVersion(); }
