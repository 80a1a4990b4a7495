# Builds the in-tree CUDA library (sm_100a only) behind include/pyrate_b200.h.
NVCC ?= nvcc
ARCH := -gencode arch=compute_100a,code=sm_100a
NVFLAGS := -O3 -std=c++17 $(ARCH) -lineinfo -Xcompiler -fPIC -Xptxas -v
SRC := pyrate_b200/csrc/pyr_trace.cu pyrate_b200/csrc/pyr_aniso.cu pyrate_b200/csrc/pyr_host.cu \
       pyrate_b200/csrc/pyr_grin_lockstep.cu pyrate_b200/csrc/pyr_host_crystal.cu
HDR := $(wildcard pyrate_b200/csrc/*.cuh) include/pyrate_b200.h
OUT := pyrate_b200/_lib/libpyrate_b200.so

all: $(OUT)

$(OUT): $(SRC) $(HDR)
	@mkdir -p pyrate_b200/_lib
	$(NVCC) $(NVFLAGS) -shared -o $@ $(SRC) 2> pyrate_b200/_lib/ptxas.log || (cat pyrate_b200/_lib/ptxas.log; exit 1)
	@grep -E "registers|spill|error|warning" pyrate_b200/_lib/ptxas.log | grep -v "^$$" | head -60 || true

# measurement build for tools/ (A/B knobs PYR_DEBUG_RECORD_LAST / PYR_LEAN_VARIANT compiled in);
# never loaded by the package unless a tool calls _native.use_tools_library()
TOOLS_OUT ?= pyrate_b200/_lib/libpyrate_b200_tools.so
tools: $(TOOLS_OUT)
$(TOOLS_OUT): $(SRC) $(HDR)
	@mkdir -p pyrate_b200/_lib
	$(NVCC) $(NVFLAGS) -DPYR_TOOLS $(TOOLS_DEFS) -shared -o $@ $(SRC) 2> pyrate_b200/_lib/ptxas_tools.log || (cat pyrate_b200/_lib/ptxas_tools.log; exit 1)

clean:
	rm -rf pyrate_b200/_lib

.PHONY: all clean tools
