# Builds libmpqc_t_cuda.so without Python (same command as `python -m mpqc_b200.build`) and the plain-C hosts.
NVCC      ?= /usr/local/cuda/bin/nvcc
NVCCFLAGS ?= -std=c++17 -O3 -lineinfo -gencode arch=compute_100a,code=sm_100a -shared -Xcompiler -fPIC -Xcompiler -Wno-format-truncation
LIB        = mpqc_b200/libmpqc_t_cuda.so
CSRC       = $(wildcard mpqc_b200/csrc/*.cu mpqc_b200/csrc/*.cuh) include/mpqc_t.h

all: $(LIB) examples/c_host examples/c_host_comm

$(LIB): $(CSRC)
	$(NVCC) $(NVCCFLAGS) -o $@ mpqc_b200/csrc/mpqc_t.cu

examples/%: examples/%.c $(LIB)
	$(CC) -std=c99 -Wall -Werror -Iinclude $< -o $@ -Lmpqc_b200 -lmpqc_t_cuda -Wl,-rpath,$(abspath mpqc_b200)

oracle:
	$(MAKE) -C oracle

clean:
	rm -f $(LIB) examples/c_host examples/c_host_comm

.PHONY: all oracle clean
