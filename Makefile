# Top-level build: the sm_100a CUDA library (C ABI), the C++ host mirror + CLI, the oracle.
#   make lib     -> matrix_solver_b200/lib/libms_b200.so
#   make cli     -> matrix_solver_b200/bin/test_mat
#   make oracle  -> oracle/_build/liboracle.so (+ oracle/_ref when /root/reference exists)
NVCC ?= /usr/local/cuda/bin/nvcc
CXX ?= g++
ARCH := -gencode arch=compute_100a,code=sm_100a
NVFLAGS := $(ARCH) -O3 -std=c++17 -lineinfo -Xcompiler -fPIC -Xcompiler -Wall -Xptxas -v $(EXTRA_NVFLAGS)
CSRC := matrix_solver_b200/csrc
HOST := matrix_solver_b200/host
LIBDIR := matrix_solver_b200/lib
BINDIR := matrix_solver_b200/bin
CU := $(CSRC)/ms_api.cu $(CSRC)/ms_kernels_util.cu $(CSRC)/ms_kernels_diag.cu \
      $(CSRC)/ms_kernels_tile.cu $(CSRC)/ms_kernels_solve.cu $(CSRC)/ms_nccl.cu $(CSRC)/ms_p2p.cu
OBJ := $(patsubst $(CSRC)/%.cu,build/%.o,$(CU))
HDR := $(CSRC)/ms_common.cuh $(CSRC)/ms_nccl.h $(CSRC)/ms_p2p.h include/ms_b200.h

.PHONY: all lib cli oracle clean
all: lib cli oracle

lib: $(LIBDIR)/libms_b200.so

build/%.o: $(CSRC)/%.cu $(HDR)
	@mkdir -p build
	$(NVCC) $(NVFLAGS) -c -o $@ $< 2> build/$*.ptxas.log || (cat build/$*.ptxas.log; false)

$(LIBDIR)/libms_b200.so: $(OBJ)
	@mkdir -p $(LIBDIR)
	$(NVCC) $(ARCH) -shared -o $@ $(OBJ) -lcudart -ldl

cli: $(BINDIR)/test_mat

$(BINDIR)/test_mat: lib $(wildcard $(HOST)/*.cpp) $(wildcard $(HOST)/*.h)
	@mkdir -p $(BINDIR)
	$(CXX) -std=c++17 -O2 -Wall -Wextra -I$(HOST) -Iinclude -o $@ $(wildcard $(HOST)/*.cpp) \
	    -L$(LIBDIR) -lms_b200 -Wl,-rpath,'$$ORIGIN/../lib' -ldl -pthread

oracle:
	$(MAKE) -C oracle oracle
	@if [ -d /root/reference/src ]; then $(MAKE) -C oracle ref; fi

clean:
	rm -rf build $(LIBDIR) $(BINDIR)
