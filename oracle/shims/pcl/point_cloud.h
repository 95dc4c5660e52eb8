#pragma once  // stand-in: the functors include it but use no PCL point cloud
