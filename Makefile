# Builds the C-ABI library (sm_100a only) and the oracle's C restatement.
NVCC      ?= nvcc
NVCCFLAGS ?= -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC
CSRC      := eigentrajectory_b200/csrc
OBJDIR    := build
SRCS      := $(wildcard $(CSRC)/*.cu)
OBJS      := $(patsubst $(CSRC)/%.cu,$(OBJDIR)/%.o,$(SRCS))
LIB       := eigentrajectory_b200/libet_b200.so

ORACLE_LIB := oracle/_build/libet_oracle.so

all: $(LIB) $(ORACLE_LIB) oracle-ref

# plain-C restatement of the bit-exact part of the oracle (test infrastructure; never linked into the product)
$(ORACLE_LIB): oracle/et_oracle_kmeans.c
	@mkdir -p oracle/_build
	gcc -O2 -ffp-contract=off -fno-fast-math -shared -fPIC -o $@ $< -lm

oracle: $(ORACLE_LIB)

# the unmodified reference staged for the GPU box (no-op where /root/reference does not exist: the prebuilt copy is used)
REFERENCE ?= /root/reference
oracle-ref:
	@if [ -d $(REFERENCE)/EigenTrajectory ]; then oracle/make_ref.sh $(REFERENCE) > /dev/null; fi

$(OBJDIR)/%.o: $(CSRC)/%.cu $(wildcard $(CSRC)/*.cuh) include/et_b200.h
	@mkdir -p $(OBJDIR)
	$(NVCC) $(NVCCFLAGS) -c $< -o $@

$(LIB): $(OBJS)
	$(NVCC) -shared -cudart static -o $@ $(OBJS) -ldl

# a plain-C host of the library (no Python, no torch): build/c_host
CUDA_HOME ?= /usr/local/cuda
example: $(LIB) examples/c_host.c
	@mkdir -p $(OBJDIR)
	gcc -std=c99 -Wall -Iinclude -I$(CUDA_HOME)/include examples/c_host.c -o $(OBJDIR)/c_host \
	    -Leigentrajectory_b200 -let_b200 -L$(CUDA_HOME)/lib64 -lcudart -lm -Wl,-rpath,$(abspath eigentrajectory_b200)

clean:
	rm -rf $(OBJDIR) $(LIB) oracle/_build oracle/_ref

.PHONY: all clean oracle oracle-ref example
