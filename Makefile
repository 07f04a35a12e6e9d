# Builds herald_b200/lib/libherald_b200.so (CUDA, sm_100a only) and the CPU oracle port.
NVCC     ?= /usr/local/cuda/bin/nvcc
ARCH     := -gencode arch=compute_100a,code=sm_100a
NVFLAGS  := $(ARCH) -O3 -std=c++17 -lineinfo -Xcompiler -fPIC,-Wall,-Wno-unused-function \
            --expt-relaxed-constexpr --expt-extended-lambda -Iinclude
# nvcc's host compiler: the image exports CXX=/opt/gcc/bin/g++ (no libgomp spec); use the distro one
CCBIN    := $(shell command -v /usr/bin/g++ || echo g++)
SRC      := $(wildcard herald_b200/csrc/*.cu)
OBJ      := $(patsubst herald_b200/csrc/%.cu,build/%.o,$(SRC))
HDR      := $(wildcard herald_b200/csrc/*.cuh) include/herald_b200.h
LIB      := herald_b200/lib/libherald_b200.so

.PHONY: all lib oracle clean
all: lib oracle

lib: $(LIB)

build/%.o: herald_b200/csrc/%.cu $(HDR)
	@mkdir -p build
	$(NVCC) -ccbin $(CCBIN) $(NVFLAGS) -c $< -o $@

$(LIB): $(OBJ)
	@mkdir -p herald_b200/lib
	$(NVCC) -ccbin $(CCBIN) $(ARCH) -shared -o $@ $^ -lcudart -ldl

oracle:
	$(MAKE) -C oracle port
	@if [ -d /root/reference ]; then $(MAKE) -C oracle ref; fi

clean:
	rm -rf build herald_b200/lib
