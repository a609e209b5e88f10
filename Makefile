# The same two build steps as `python -m bonsai_b200.build`, for C++ users: the sm_100a library and the `bonsai` CLI, in-tree.
NVCC ?= /usr/local/cuda/bin/nvcc
HOSTCXX ?= /usr/bin/g++
PKG  := bonsai_b200
CSRC := $(PKG)/csrc
LIB  := $(PKG)/libbonsai_b200.so
CLI  := $(PKG)/bin/bonsai
NVCCFLAGS := -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC,-ffp-contract=off,-Wall,-pthread -shared -ldl

all: $(LIB) $(CLI)

$(LIB): $(CSRC)/bns_kernels.cu $(CSRC)/bns_api.cu $(CSRC)/bns_pack.cpp $(CSRC)/bns_pack.h $(CSRC)/bns_device.cuh $(CSRC)/bns_classify_u.cuh $(CSRC)/bns_kernels.h $(CSRC)/bns_host_util.h include/bonsai_b200.h
	$(NVCC) $(NVCCFLAGS) -o $@ $(CSRC)/bns_kernels.cu $(CSRC)/bns_api.cu $(CSRC)/bns_pack.cpp

$(CLI): $(CSRC)/cli/bonsai_main.cpp include/bonsai_b200/bonsai.hpp include/bonsai_b200.h $(LIB)
	mkdir -p $(PKG)/bin
	$(HOSTCXX) -O2 -std=c++17 -Wall -o $@ $(CSRC)/cli/bonsai_main.cpp -L$(PKG) -lbonsai_b200 -lz -lpthread '-Wl,-rpath,$$ORIGIN/..'

# the CPU checkers (test infrastructure): the C restatement, and the driver over the unmodified reference headers where they exist
oracle:
	$(MAKE) -C oracle

test-cpu: all
	python -m pytest tests -q -m "not gpu"

test-gpu: all
	python -m pytest tests -q -m gpu

clean:
	rm -f $(LIB) $(CLI)

.PHONY: all oracle test-cpu test-gpu clean
