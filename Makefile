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
DEPS      := $(OBJS:.o=.d)

all: $(LIB) oracle

# per-object header dependencies (nvcc -MMD): a header edit rebuilds only the units that include it
$(OBJDIR)/%.o: $(CSRC)/%.cu
	@mkdir -p $(OBJDIR)
	$(NVCC) $(NVFLAGS) $(EXTRA) -MMD -MP -MF $(OBJDIR)/$*.d -c $< -o $@

-include $(DEPS)

$(LIB): $(OBJS)
	$(NVCC) $(ARCH) -shared -cudart shared -o $@ $(OBJS)

oracle:
	$(MAKE) -C oracle

clean:
	rm -rf build $(LIB)
	$(MAKE) -C oracle clean

.PHONY: all oracle clean
