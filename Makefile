# Builds the sm_100a shared library (C ABI in include/sdm_b200.h) and the GPU-side tools.
NVCC      ?= nvcc
PKG       := simple-diffusion-model_b200
CSRC      := $(PKG)/csrc
BUILD     := build
LIB       := $(PKG)/lib/libsdm_b200.so
NVFLAGS   := -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -I$(CSRC) -Iinclude -Xcompiler -fPIC
SRCS      := $(wildcard $(CSRC)/*.cu)
OBJS      := $(patsubst $(CSRC)/%.cu,$(BUILD)/%.o,$(SRCS))

all: $(LIB) tools

$(BUILD)/%.o: $(CSRC)/%.cu $(wildcard $(CSRC)/*.h) $(wildcard $(CSRC)/*.cuh) include/sdm_b200.h
	@mkdir -p $(BUILD)
	$(NVCC) $(NVFLAGS) -c $< -o $@

$(LIB): $(OBJS)
	@mkdir -p $(PKG)/lib
	$(NVCC) -shared -o $@ $(OBJS) -lcudart

tools: $(BUILD)/test_igemm

$(BUILD)/test_igemm: tools/test_igemm.cu $(LIB)
	$(NVCC) $(NVFLAGS) -o $@ $< -L$(PKG)/lib -lsdm_b200 -Xlinker -rpath -Xlinker '$$ORIGIN/../$(PKG)/lib'

oracle:
	@true

clean:
	rm -rf $(BUILD) $(LIB)

.PHONY: all tools clean oracle
