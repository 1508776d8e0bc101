#pragma once
#include <Core/Geometry/PointCloud.h>
#include <Core/Registration/Registration.h>
#include <Core/Registration/TransformationEstimation.h>
