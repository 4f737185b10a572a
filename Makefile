# Top-level build: libbgsa_b200.so (CUDA kernels + C ABI, sm_100a only), the C host tools
# (aligner, convert), the drop-in align_core shims and the test-only host simulator.
NVCC     ?= nvcc
GCC      ?= gcc
ARCH     := -gencode arch=compute_100a,code=sm_100a
# EXTRA: additional nvcc flags for experiment builds, e.g. make lib BUILD=build/x LIB=bgsa_b200/libx.so EXTRA=-DBGSA_FMA_SHIFT=0
EXTRA    ?=
NVFLAGS  := $(ARCH) -O3 -lineinfo -std=c++17 -Xcompiler -fPIC -Xcompiler -Wall -Xcompiler -Wno-unused-function -Xcompiler -Wno-unknown-pragmas $(EXTRA)
CSRC     := bgsa_b200/csrc
HOST     := bgsa_b200/host
BUILD    ?= build
LIB      ?= bgsa_b200/libbgsa_b200.so

HDRS := $(wildcard $(CSRC)/*.cuh $(CSRC)/*.h) include/bgsa_b200.h

OBJS := $(BUILD)/api.o $(BUILD)/inst_misc.o $(BUILD)/inst_myers_g.o $(BUILD)/inst_myers_s.o \
        $(BUILD)/inst_bp_p0.o $(BUILD)/inst_bp_p1.o $(BUILD)/inst_bp_p2.o \
        $(BUILD)/inst_bp_n0.o $(BUILD)/inst_bp_n1.o $(BUILD)/inst_bp_n2.o \
        $(BUILD)/inst_bp_s0.o $(BUILD)/inst_bp_s1.o $(BUILD)/inst_bp_s2.o

.PHONY: all lib tools sim clean
all: lib tools sim

lib: $(LIB)
$(LIB): $(OBJS)
	$(NVCC) $(ARCH) -shared -o $@ $(OBJS) -lcudart

$(BUILD)/api.o: $(CSRC)/api.cu $(HDRS) | $(BUILD)
	$(NVCC) $(NVFLAGS) -c $< -o $@
$(BUILD)/inst_misc.o: $(CSRC)/inst_misc.cu $(HDRS) | $(BUILD)
	$(NVCC) $(NVFLAGS) -c $< -o $@
$(BUILD)/inst_myers_g.o: $(CSRC)/inst_myers.cu $(HDRS) | $(BUILD)
	$(NVCC) $(NVFLAGS) -DBGSA_MYERS_MODE=0 -c $< -o $@
$(BUILD)/inst_myers_s.o: $(CSRC)/inst_myers.cu $(HDRS) | $(BUILD)
	$(NVCC) $(NVFLAGS) -DBGSA_MYERS_MODE=1 -c $< -o $@
# BitPAl: one object per (scheme id, variant: p packed, n non-packed, s packed semi-global) -- keep in sync with BGSA_SCHEMES in instances.h
$(BUILD)/inst_bp_p0.o: $(CSRC)/inst_bitpal.cu $(HDRS) | $(BUILD)
	$(NVCC) $(NVFLAGS) -DBGSA_SCHEME_ID=0 -DBGSA_M=2 -DBGSA_I=-3 -DBGSA_G=-5 -DBGSA_PACKED=1 -c $< -o $@
$(BUILD)/inst_bp_p1.o: $(CSRC)/inst_bitpal.cu $(HDRS) | $(BUILD)
	$(NVCC) $(NVFLAGS) -DBGSA_SCHEME_ID=1 -DBGSA_M=1 -DBGSA_I=-1 -DBGSA_G=-1 -DBGSA_PACKED=1 -c $< -o $@
$(BUILD)/inst_bp_p2.o: $(CSRC)/inst_bitpal.cu $(HDRS) | $(BUILD)
	$(NVCC) $(NVFLAGS) -DBGSA_SCHEME_ID=2 -DBGSA_M=1 -DBGSA_I=-3 -DBGSA_G=-2 -DBGSA_PACKED=1 -c $< -o $@
$(BUILD)/inst_bp_n0.o: $(CSRC)/inst_bitpal.cu $(HDRS) | $(BUILD)
	$(NVCC) $(NVFLAGS) -DBGSA_SCHEME_ID=0 -DBGSA_M=2 -DBGSA_I=-3 -DBGSA_G=-5 -DBGSA_PACKED=0 -c $< -o $@
$(BUILD)/inst_bp_n1.o: $(CSRC)/inst_bitpal.cu $(HDRS) | $(BUILD)
	$(NVCC) $(NVFLAGS) -DBGSA_SCHEME_ID=1 -DBGSA_M=1 -DBGSA_I=-1 -DBGSA_G=-1 -DBGSA_PACKED=0 -c $< -o $@
$(BUILD)/inst_bp_n2.o: $(CSRC)/inst_bitpal.cu $(HDRS) | $(BUILD)
	$(NVCC) $(NVFLAGS) -DBGSA_SCHEME_ID=2 -DBGSA_M=1 -DBGSA_I=-3 -DBGSA_G=-2 -DBGSA_PACKED=0 -c $< -o $@

$(BUILD)/inst_bp_s0.o: $(CSRC)/inst_bitpal.cu $(HDRS) | $(BUILD)
	$(NVCC) $(NVFLAGS) -DBGSA_SCHEME_ID=0 -DBGSA_M=2 -DBGSA_I=-3 -DBGSA_G=-5 -DBGSA_PACKED=2 -c $< -o $@
$(BUILD)/inst_bp_s1.o: $(CSRC)/inst_bitpal.cu $(HDRS) | $(BUILD)
	$(NVCC) $(NVFLAGS) -DBGSA_SCHEME_ID=1 -DBGSA_M=1 -DBGSA_I=-1 -DBGSA_G=-1 -DBGSA_PACKED=2 -c $< -o $@
$(BUILD)/inst_bp_s2.o: $(CSRC)/inst_bitpal.cu $(HDRS) | $(BUILD)
	$(NVCC) $(NVFLAGS) -DBGSA_SCHEME_ID=2 -DBGSA_M=1 -DBGSA_I=-3 -DBGSA_G=-2 -DBGSA_PACKED=2 -c $< -o $@

# test-only: the DP column functions compiled for the HOST (no GPU needed), see tests/host_sim.cu
sim: tests/libhost_sim.so
tests/libhost_sim.so: tests/host_sim.cu $(HDRS)
	$(NVCC) -O2 -std=c++17 -Xcompiler -fPIC -shared -I$(CSRC) -o $@ $<

SHIMS := myers_cpu semiglobal_cpu banded_cpu myers_sse bitpal_avx2 bitpal_avx512
SHIM_LIBS := $(foreach v,$(SHIMS),bgsa_b200/libalign_core_$(v).so)

tools: bgsa_b200/aligner bgsa_b200/convert $(SHIM_LIBS)

# host code is plain C (gcc); it reaches CUDA only through the C ABI of libbgsa_b200.so
bgsa_b200/aligner: $(HOST)/aligner.c include/bgsa_b200.h $(LIB)
	$(GCC) -O2 -Wall -o $@ $(HOST)/aligner.c -Lbgsa_b200 -lbgsa_b200 -Wl,-rpath,'$$ORIGIN'

# no GPU code: FASTA/FASTQ -> line format, result file -> text (the reference's convert tool)
bgsa_b200/convert: $(HOST)/convert.c
	$(GCC) -O2 -Wall -o $@ $(HOST)/convert.c

upper = $(shell echo $(1) | tr a-z A-Z)
bgsa_b200/libalign_core_%.so: $(HOST)/align_core_shim.c include/align_core.h include/bgsa_b200.h $(LIB)
	$(GCC) -O2 -Wall -shared -fPIC -DBGSA_SHIM_$(call upper,$*) -o $@ $(HOST)/align_core_shim.c -Lbgsa_b200 -lbgsa_b200 -Wl,-rpath,'$$ORIGIN'

$(BUILD):
	mkdir -p $(BUILD)

clean:
	rm -rf $(BUILD) $(LIB) tests/libhost_sim.so bgsa_b200/aligner bgsa_b200/convert $(SHIM_LIBS)
