# Builds the product library (CUDA, sm_100a only) and the test oracle (plain C).
#   make            -> quiver_b200/lib/libquivergpu.so, quiver_b200/lib/libquiverhost.so, oracle/liboracle.so
#   make -j8 lib    -> only the CUDA library
NVCC      ?= /usr/local/cuda/bin/nvcc
CC        ?= gcc
CXX       ?= g++
ARCH      := -gencode arch=compute_100a,code=sm_100a
NVFLAGS   := -std=c++17 -O3 $(ARCH) -lineinfo -Xcompiler -fPIC -Xcompiler -Wall -Xcompiler -Wno-unused-function \
             --expt-relaxed-constexpr -Xptxas -warn-spills
# make TC_INSTRUMENT=1: per-role cycle counters inside the tensor-core scan (tools/tc_timing.py)
ifeq ($(TC_INSTRUMENT),1)
NVFLAGS   += -DQG_TC_INSTRUMENT
endif
# make EXTRA_NVFLAGS=-D...: experiment switches of single kernels
NVFLAGS   += $(EXTRA_NVFLAGS)
CSRC      := quiver_b200/csrc
BUILD     := build/csrc
LIBDIR    := quiver_b200/lib

CU_SRCS   := $(wildcard $(CSRC)/*.cu)
CU_OBJS   := $(patsubst $(CSRC)/%.cu,$(BUILD)/%.o,$(CU_SRCS))
CU_HDRS   := $(wildcard $(CSRC)/*.cuh) include/quiver_gpu.h

HOST_SRCS := $(wildcard quiver_b200/host/*.cpp)
HOST_HDRS := $(wildcard quiver_b200/host/*.hpp) include/quiver_gpu.h include/quiver_host.h

ifeq ($(strip $(HOST_SRCS)),)
all: lib oracle
else
all: lib host oracle
endif

lib: $(LIBDIR)/libquivergpu.so
host: $(LIBDIR)/libquiverhost.so
oracle: oracle/liboracle.so

$(BUILD)/%.o: $(CSRC)/%.cu $(CU_HDRS)
	@mkdir -p $(BUILD)
	$(NVCC) $(NVFLAGS) -c $< -o $@

$(LIBDIR)/libquivergpu.so: $(CU_OBJS)
	@mkdir -p $(LIBDIR)
	$(NVCC) $(ARCH) -shared -cudart static -o $@ $(CU_OBJS) -ldl

# Host side above the C ABI (C++ because the reference host is compiled Go and no Go toolchain
# exists in this image). -ffp-contract=off: the rerank arithmetic must not fuse (hybrid_index.go:552).
$(LIBDIR)/libquiverhost.so: $(HOST_SRCS) $(HOST_HDRS) $(LIBDIR)/libquivergpu.so
	@mkdir -p $(LIBDIR)
	$(CXX) -std=c++17 -O2 -fPIC -shared -pthread -ffp-contract=off -Wall -Iinclude -o $@ $(HOST_SRCS) \
	    -L$(LIBDIR) -lquivergpu -Wl,-rpath,'$$ORIGIN'

# The oracle is test infrastructure: strict IEEE, no contraction (Go/amd64 never fuses).
ORACLE_SRCS := $(wildcard oracle/*.c)
oracle/liboracle.so: $(ORACLE_SRCS) oracle/synth.h
	$(CC) -O2 -ffp-contract=off -fno-fast-math -fPIC -shared -Wall -o $@ $(ORACLE_SRCS) -lm -lpthread

clean:
	rm -rf build $(LIBDIR)/*.so oracle/liboracle.so

.PHONY: all lib host oracle clean
