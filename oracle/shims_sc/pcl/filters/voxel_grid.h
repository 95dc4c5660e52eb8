// Empty stand-in: Scancontext.h includes it, the ScanContext code uses nothing from it.  TEST INFRASTRUCTURE (oracle/_ref/libref_scancontext.so).
#pragma once
