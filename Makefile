# libstorm_b200.so for a C / C++ caller without Python: `make` (the same commands stormbitmaps_b200/build.py runs).
#   make            stormbitmaps_b200/libstorm_b200.so      (nvcc, sm_100a; cross-compiles without a GPU)
#   make dropin     tests/drivers/dropin_driver.c linked against it  -> build/dropin_driver
#   make clean
NVCC    ?= nvcc
CC      ?= gcc
PKG     := stormbitmaps_b200
CSRC    := $(PKG)/csrc
OBJDIR  := $(PKG)/_obj
LIB     := $(PKG)/libstorm_b200.so
ARCH    := -gencode arch=compute_100a,code=sm_100a
NVFLAGS := $(ARCH) -lineinfo -O3 -std=c++17 -Xcompiler -fPIC -I include -I $(CSRC)
SRCS    := $(sort $(wildcard $(CSRC)/*.cu))
OBJS    := $(patsubst $(CSRC)/%.cu,$(OBJDIR)/%.o,$(SRCS))
HDRS    := $(wildcard $(CSRC)/*.cuh) $(wildcard $(CSRC)/*.h) $(wildcard include/*.h)

all: $(LIB)

$(OBJDIR)/%.o: $(CSRC)/%.cu $(HDRS)
	@mkdir -p $(OBJDIR)
	$(NVCC) $(NVFLAGS) -c $< -o $@

$(LIB): $(OBJS)
	$(NVCC) -shared -o $@ $(OBJS) $(ARCH)

# the reference-facing C99 caller (storm.h only), the way a maintainer's program links
dropin: $(LIB)
	@mkdir -p build
	$(CC) -std=c99 -O2 -Wall -I include tests/drivers/dropin_driver.c -L $(PKG) -lstorm_b200 -Wl,-rpath,'$$ORIGIN/../$(PKG)' -o build/dropin_driver

clean:
	rm -rf $(OBJDIR) $(LIB) build

.PHONY: all dropin clean
