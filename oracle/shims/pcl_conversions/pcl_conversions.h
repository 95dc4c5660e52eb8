#pragma once  // stand-in: the functors include it but convert nothing
