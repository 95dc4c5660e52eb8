# Builds for a C++ consumer without Python (the Python entry point `__graft_entry__.build()` does the same in-tree):
#   make lib        libilsm_cuda.so (sm_100a) from intensity_based_lidar_slam_for_me-_b200/csrc/*.cu
#   make oracle     the CPU oracle (test infrastructure) and, when /root/reference exists, oracle/_ref
#   make cpp-test   tests/cpp/host_mirror_test (needs lib + oracle; run it on a B200)
NVCC     ?= nvcc
CXX      ?= g++
PKG      := intensity_based_lidar_slam_for_me-_b200
CSRC     := $(PKG)/csrc
OBJDIR   := $(PKG)/build
LIB      := $(PKG)/libilsm_cuda.so
ARCH     := -gencode arch=compute_100a,code=sm_100a
NVFLAGS  := -O3 -std=c++17 $(ARCH) -lineinfo -Xcompiler -fPIC,-fvisibility=hidden
SRCS     := $(wildcard $(CSRC)/*.cu)
OBJS     := $(patsubst $(CSRC)/%.cu,$(OBJDIR)/%.o,$(SRCS))
HDRS     := $(wildcard $(CSRC)/*.cuh $(CSRC)/*.hpp) include/ilsm.h

lib: $(LIB)

# Files whose float results are compared bit for bit with the reference's x86-64 (no FMA) build are compiled without
# a*b+c contraction; registration.cu keeps its bit-exact parts in explicit round-to-nearest intrinsics (see _build.py).
$(OBJDIR)/registration.o: FMAD := true
FMAD ?= false

$(OBJDIR)/%.o: $(CSRC)/%.cu $(HDRS)
	@mkdir -p $(OBJDIR)
	$(NVCC) $(NVFLAGS) -fmad=$(FMAD) -c $< -o $@

$(LIB): $(OBJS)
	$(NVCC) -shared $(ARCH) -o $@ $(OBJS)

oracle:
	$(MAKE) -C oracle

cpp-test: lib oracle
	$(CXX) -std=c++14 -O2 -ffp-contract=off -Wall -Iinclude tests/cpp/host_mirror_test.cpp -o tests/cpp/host_mirror_test \
	  -L$(PKG) -lilsm_cuda -Loracle/_build -lilsm_oracle -Wl,-rpath,$(abspath $(PKG)) -Wl,-rpath,$(abspath oracle/_build)

clean:
	rm -rf $(OBJDIR) $(LIB) $(LIB).stamp tests/cpp/host_mirror_test

.PHONY: lib oracle cpp-test clean
