# Builds the in-tree CUDA library (sm_100a only) behind include/pyrate_b200.h.
NVCC ?= nvcc
ARCH := -gencode arch=compute_100a,code=sm_100a
NVFLAGS := -O3 -std=c++17 $(ARCH) -lineinfo -Xcompiler -fPIC -Xptxas -v
SRC := pyrate_b200/csrc/pyr_trace.cu pyrate_b200/csrc/pyr_aniso.cu pyrate_b200/csrc/pyr_host.cu
HDR := $(wildcard pyrate_b200/csrc/*.cuh) include/pyrate_b200.h
OUT := pyrate_b200/_lib/libpyrate_b200.so

all: $(OUT)

$(OUT): $(SRC) $(HDR)
	@mkdir -p pyrate_b200/_lib
	$(NVCC) $(NVFLAGS) -shared -o $@ $(SRC) 2> pyrate_b200/_lib/ptxas.log || (cat pyrate_b200/_lib/ptxas.log; exit 1)
	@grep -E "registers|spill|error|warning" pyrate_b200/_lib/ptxas.log | grep -v "^$$" | head -60 || true

clean:
	rm -rf pyrate_b200/_lib

.PHONY: all clean
