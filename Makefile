# libvcfb200.so -- sm_100a only.  `make` builds in-tree (vcf_b200/libvcfb200.so) so the
# shared object travels with the repo snapshot to the GPU box.
NVCC      ?= nvcc
ARCH      := -gencode arch=compute_100a,code=sm_100a
NVCCFLAGS := $(ARCH) -O3 -std=c++17 -lineinfo -Xcompiler -fPIC -Iinclude -Ivcf_b200/csrc $(EXTRA)
SRCS      := vcf_b200/csrc/api.cu vcf_b200/csrc/kernels_general.cu vcf_b200/csrc/kernels_anyb.cu vcf_b200/csrc/kernels_rd.cu vcf_b200/csrc/kernels_tile.cu vcf_b200/csrc/kernels_fast.cu vcf_b200/csrc/kernels_packed.cu vcf_b200/csrc/kernels_b16.cu vcf_b200/csrc/kernels_b16f.cu vcf_b200/csrc/kernels_dec2t.cu vcf_b200/csrc/kernels_dec32.cu vcf_b200/csrc/kernels_tc.cu vcf_b200/csrc/kernels_color.cu vcf_b200/csrc/kernels_stats.cu vcf_b200/csrc/kernels_motion.cu vcf_b200/csrc/kernels_deflate.cu
SRCS      := $(wildcard $(SRCS))
OBJS      := $(patsubst vcf_b200/csrc/%.cu,build/%.o,$(SRCS))
HDRS      := $(wildcard vcf_b200/csrc/*.cuh) include/vcfb200.h
LIB       := vcf_b200/libvcfb200.so

all: $(LIB)

# the packed (f32x2) kernels need contraction off: ptxas fuses explicit .rn packed ops otherwise
build/kernels_packed.o build/kernels_b16.o: NVCCFLAGS += -fmad=false

build/%.o: vcf_b200/csrc/%.cu $(HDRS)
	@mkdir -p build
	$(NVCC) $(NVCCFLAGS) -Xptxas -v -c $< -o $@ 2> build/$*.ptxas.log || (cat build/$*.ptxas.log; exit 1)

$(LIB): $(OBJS)
	$(NVCC) $(ARCH) -shared -o $@ $(OBJS)

codelets:
	python vcf_b200/codegen/gen_cuda.py

clean:
	rm -rf build $(LIB)

.PHONY: all clean codelets
