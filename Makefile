# Builds the C-ABI shared library (sm_100a only) and the CPU oracle.
NVCC      ?= nvcc
ARCH      := -gencode arch=compute_100a,code=sm_100a
# -cudart shared: the library links libcudart.so instead of carrying a private copy of the runtime
NVFLAGS   := -O3 -std=c++17 -lineinfo $(ARCH) -Xcompiler -fPIC --expt-relaxed-constexpr -cudart shared
CSRC      := ode_uncertainty_b200/csrc
OBJDIR    := build/obj
LIB       := ode_uncertainty_b200/libodeu.so
SRCS      := $(wildcard $(CSRC)/*.cu)
OBJS      := $(patsubst $(CSRC)/%.cu,$(OBJDIR)/%.o,$(SRCS))
HDRS      := $(wildcard $(CSRC)/*.cuh) $(wildcard $(CSRC)/*.h) include/odeu.h

all: $(LIB) oracle

$(OBJDIR)/%.o: $(CSRC)/%.cu $(HDRS)
	@mkdir -p $(OBJDIR)
	$(NVCC) $(NVFLAGS) $(EXTRA) -c $< -o $@

$(LIB): $(OBJS)
	$(NVCC) $(ARCH) -shared -cudart shared -o $@ $(OBJS)

oracle:
	$(MAKE) -C oracle

clean:
	rm -rf build $(LIB)
	$(MAKE) -C oracle clean

.PHONY: all oracle clean
