#pragma once  // stand-in: nothing of ceres/rotation.h is used by the functors
