# Build of the B200-native hot-path library.  sm_100a only (tcgen05 / TMEM / bulk-copy engine).
NVCC      ?= /usr/local/cuda/bin/nvcc
ARCH      := -gencode arch=compute_100a,code=sm_100a
NVCCFLAGS := $(ARCH) -O3 -std=c++17 -lineinfo -Xcompiler -fPIC,-Wall,-Wno-unused-function -Xptxas -v
CSRC      := monocularsfm_b200/csrc
LIB       := monocularsfm_b200/libmsfm_b200.so
OBJDIR    := build/obj

CU_SRCS   := $(wildcard $(CSRC)/*.cu)
CU_OBJS   := $(patsubst $(CSRC)/%.cu,$(OBJDIR)/%.o,$(CU_SRCS))
HDRS      := $(wildcard $(CSRC)/*.cuh) $(wildcard $(CSRC)/*.hpp) include/msfm_b200.h

HOSTDIR   := monocularsfm_b200/host
HOSTLIB   := monocularsfm_b200/libmsfm_host.so
HOST_SRCS := $(wildcard $(HOSTDIR)/src/*.cpp)
HOST_HDRS := $(wildcard $(HOSTDIR)/include/*/*.h) $(wildcard $(HOSTDIR)/src/*.h) include/msfm_b200.h
CXX       ?= g++
CXXFLAGS  := -O2 -std=c++14 -fPIC -Wall -Wno-sign-compare -I$(HOSTDIR)/include -Iinclude

all: $(LIB) $(HOSTLIB) build/host_test

$(OBJDIR)/%.o: $(CSRC)/%.cu $(HDRS)
	@mkdir -p $(OBJDIR)
	$(NVCC) $(NVCCFLAGS) -c $< -o $@ 2> $(OBJDIR)/$*.ptxas.log || (cat $(OBJDIR)/$*.ptxas.log; exit 1)

$(LIB): $(CU_OBJS)
	$(NVCC) $(ARCH) -shared -o $@ $^ -lcusolver -lcublas -ldl -Xlinker -rpath=/usr/local/cuda/lib64

# C++ classes with the reference's names (FeatureUtils, FeatureMatcher, Database, BundleData, CeresBundelOptimizer)
$(HOSTLIB): $(HOST_SRCS) $(HOST_HDRS) $(LIB)
	$(CXX) $(CXXFLAGS) -shared -o $@ $(HOST_SRCS) -Lmonocularsfm_b200 -lmsfm_b200 -ldl -Wl,-rpath,'$$ORIGIN'

build/host_test: $(HOSTDIR)/test/host_test.cpp $(HOSTLIB)
	@mkdir -p build
	$(CXX) $(CXXFLAGS) -o $@ $< -Lmonocularsfm_b200 -lmsfm_host -lmsfm_b200 -Wl,-rpath,'$$ORIGIN/../monocularsfm_b200'

clean:
	rm -rf build/obj $(LIB) $(HOSTLIB) build/host_test

.PHONY: all clean

microbench: build/microbench
build/microbench: tools/microbench.cu $(CSRC)/ptx.cuh
	@mkdir -p build
	$(NVCC) $(ARCH) -O3 -std=c++17 -lineinfo -o $@ $<

microbench_pair: build/microbench_pair
build/microbench_pair: tools/microbench_pair.cu $(CSRC)/ptx.cuh
	@mkdir -p build
	$(NVCC) $(ARCH) -O3 -std=c++17 -lineinfo -o $@ $<

microbench_alu: build/microbench_alu
build/microbench_alu: tools/microbench_alu.cu
	@mkdir -p build
	$(NVCC) $(ARCH) -O3 -std=c++17 -lineinfo -o $@ $<

microbench_f64: build/microbench_f64
build/microbench_f64: tools/microbench_f64.cu
	@mkdir -p build
	$(NVCC) $(ARCH) -O3 -std=c++17 -lineinfo -o $@ $<
