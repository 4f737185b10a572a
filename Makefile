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

# Scoring schemes compiled into the library ("match,mismatch,gap", scheme id = position).  The reference runs its Java
# generator once per scheme (generator/.../Main.java:240-315); here a scheme is a set of template instances:
#     make SCHEMES="2,-3,-5 1,-1,-1 1,-3,-2 3,-2,-4"
# The first three are the default and are what instances.h falls back to when BGSA_SCHEMES is not defined.
SCHEMES  ?= 2,-3,-5 1,-1,-1 1,-3,-2
comma    := ,
SCHEME_IDS := $(shell i=0; for s in $(SCHEMES); do echo $$i; i=$$((i+1)); done)
scheme_of = $(word $(shell echo $$(($(1)+1))),$(SCHEMES))
SCHEME_LIST := $(foreach i,$(SCHEME_IDS),X($(i)$(comma) $(subst $(comma),$(comma) ,$(call scheme_of,$(i)))))
# (nvcc splits -D values at commas, so the list travels in a generated header)
SCHEME_HDR := $(BUILD)/bgsa_schemes.h
SCHEME_DEF := -include $(SCHEME_HDR)

BP_OBJS := $(foreach i,$(SCHEME_IDS),$(BUILD)/inst_bp_p$(i).o $(BUILD)/inst_bp_n$(i).o $(BUILD)/inst_bp_s$(i).o)
OBJS := $(BUILD)/api.o $(BUILD)/jit.o $(BUILD)/host_pack.o $(BUILD)/inst_misc.o $(BUILD)/inst_myers_g.o $(BUILD)/inst_myers_s.o $(BP_OBJS)

.PHONY: all lib tools sim clean
all: lib tools sim

lib: $(LIB)
$(LIB): $(OBJS)
	$(NVCC) $(ARCH) -shared -o $@ $(OBJS) -lcudart -lpthread -ldl

$(SCHEME_HDR): Makefile | $(BUILD)
	@echo '#define BGSA_SCHEMES(X) $(SCHEME_LIST)' > $@.tmp; cmp -s $@.tmp $@ || mv $@.tmp $@; rm -f $@.tmp
.PHONY: FORCE
$(SCHEME_HDR): FORCE

$(BUILD)/api.o: $(CSRC)/api.cu $(HDRS) $(SCHEME_HDR) | $(BUILD)
	$(NVCC) $(NVFLAGS) $(SCHEME_DEF) -c $< -o $@
# run-time instantiation of scoring schemes outside SCHEMES (NVRTC): the kernel headers travel inside the library
JIT_HDRS := $(CSRC)/bgsa_common.cuh $(CSRC)/align_kernel.cuh $(CSRC)/rows_kernel.cuh $(CSRC)/bitpal.cuh
$(BUILD)/jit_sources.inc: $(JIT_HDRS) tools/embed_sources.py | $(BUILD)
	python tools/embed_sources.py $@ $(JIT_HDRS)
$(BUILD)/jit.o: $(CSRC)/jit.cu $(HDRS) $(BUILD)/jit_sources.inc | $(BUILD)
	$(NVCC) $(NVFLAGS) -I$(BUILD) -c $< -o $@
# host-side pack (plain C++, g++; AVX2 code paths are selected at run time)
$(BUILD)/host_pack.o: $(CSRC)/host_pack.cpp $(CSRC)/host_pack.h | $(BUILD)
	g++ -O3 -std=c++17 -fPIC -Wall -c $< -o $@
$(BUILD)/inst_misc.o: $(CSRC)/inst_misc.cu $(HDRS) $(SCHEME_HDR) | $(BUILD)
	$(NVCC) $(NVFLAGS) $(SCHEME_DEF) -c $< -o $@
$(BUILD)/inst_myers_g.o: $(CSRC)/inst_myers.cu $(HDRS) | $(BUILD)
	$(NVCC) $(NVFLAGS) -DBGSA_MYERS_MODE=0 -c $< -o $@
$(BUILD)/inst_myers_s.o: $(CSRC)/inst_myers.cu $(HDRS) | $(BUILD)
	$(NVCC) $(NVFLAGS) -DBGSA_MYERS_MODE=1 -c $< -o $@
# BitPAl: one object per (scheme id, variant: p packed = 1, n non-packed = 0, s packed semi-global = 2)
define BP_RULE
$$(BUILD)/inst_bp_$(2)$(1).o: $$(CSRC)/inst_bitpal.cu $$(HDRS) | $$(BUILD)
	$$(NVCC) $$(NVFLAGS) -DBGSA_SCHEME_ID=$(1) -DBGSA_M=$(word 1,$(subst $(comma), ,$(call scheme_of,$(1)))) -DBGSA_I=$(word 2,$(subst $(comma), ,$(call scheme_of,$(1)))) -DBGSA_G=$(word 3,$(subst $(comma), ,$(call scheme_of,$(1)))) -DBGSA_PACKED=$(3) -c $$< -o $$@
endef
$(foreach i,$(SCHEME_IDS),$(eval $(call BP_RULE,$(i),p,1))$(eval $(call BP_RULE,$(i),n,0))$(eval $(call BP_RULE,$(i),s,2)))

# test-only: the DP column functions compiled for the HOST (no GPU needed), see tests/host_sim.cu
sim: tests/libhost_sim.so
tests/libhost_sim.so: tests/host_sim.cu $(HDRS) $(SCHEME_HDR)
	$(NVCC) -O2 -std=c++17 -Xcompiler -fPIC -shared -I$(CSRC) $(SCHEME_DEF) -o $@ $<

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
