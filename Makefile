# Builds libvo_b200.so (the CUDA hot path + its C ABI, sm_100a only) and the CPU oracle.
# nvcc cross-compiles without a GPU.  The .so is built IN-TREE so it travels to the GPU box.
NVCC ?= nvcc
PKG := visual-odometry_b200
CSRC := $(PKG)/csrc
LIB := $(PKG)/lib/libvo_b200.so
NVCCFLAGS := -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 \
             -Xcompiler -fPIC,-Wall,-ffp-contract=off --expt-relaxed-constexpr
SRCS := $(CSRC)/lib.cu $(CSRC)/stage.cu $(CSRC)/nn.cu $(CSRC)/nn_tc.cu $(CSRC)/comm.cu $(CSRC)/picp.cu $(CSRC)/triangulate.cu $(CSRC)/pipeline.cu
OBJS := $(SRCS:$(CSRC)/%.cu=build/%.o)
HDRS := $(wildcard $(CSRC)/*.cuh) include/vo_b200.h

APPS := $(PKG)/host/bin/nn_sharded

all: $(LIB) oracle $(APPS)

# a plain C++ caller of the multi-GPU C ABI (no Eigen, no reference sources needed)
$(PKG)/host/bin/nn_sharded: $(PKG)/host/apps/nn_sharded.cpp $(LIB) include/vo_b200.h
	@mkdir -p $(dir $@)
	$(CXX) -std=c++17 -O2 -I include $< -L $(PKG)/lib -lvo_b200 -Wl,-rpath,'$$ORIGIN/../../lib' -o $@

$(LIB): $(OBJS)
	@mkdir -p $(dir $@)
	$(NVCC) -shared -o $@ $(OBJS) -lcudart_static -lpthread -ldl -lrt

# triangulate.cu mirrors the oracle's operation order exactly, so it is compiled without FMA
# contraction (-fmad=false): its results are then bit-identical to the CPU restatement.
build/triangulate.o: EXTRA := -fmad=false
# make TC_PROFILE=1: cycle counters inside nn_tc_filter_kernel (see nn_tc.cu)
ifdef TC_PROFILE
build/nn_tc.o: EXTRA := -DNN_TC_PROFILE
endif

build/%.o: $(CSRC)/%.cu $(HDRS)
	@mkdir -p build
	$(NVCC) $(NVCCFLAGS) $(EXTRA) -Xptxas -v -c $< -o $@ 2> build/$*.ptxas.log || (cat build/$*.ptxas.log; exit 1)

oracle:
	$(MAKE) -C oracle

clean:
	rm -rf build $(LIB)
	$(MAKE) -C oracle clean
.PHONY: all oracle clean
