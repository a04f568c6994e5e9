# Build: product shared library (CUDA + host C++) and the test-only oracle.
NVCC ?= /usr/local/cuda/bin/nvcc
CSRC := cavif_rs_b200/csrc
LIB := cavif_rs_b200/libcavif_b200.so
ORACLE := oracle/liboracle.so
GENCODE := -gencode arch=compute_100a,code=sm_100a
NVFLAGS := -O3 -std=c++17 -lineinfo $(GENCODE) -Xcompiler -fPIC,-Wall,-Wno-missing-braces,-Wno-unused-function -cudart static -I include -I $(CSRC)
CU := $(wildcard $(CSRC)/*.cu)
CPP := $(wildcard $(CSRC)/*.cpp)
HDR := $(wildcard $(CSRC)/*.h) $(wildcard include/*.h)
OSRC := $(wildcard oracle/*.c)

all: lib oracle
lib: $(LIB)
oracle: $(ORACLE)

$(LIB): $(CU) $(CPP) $(HDR)
	$(NVCC) $(NVFLAGS) -shared -o $@ $(CU) $(CPP) -lpthread

$(ORACLE): $(OSRC) $(HDR) $(wildcard oracle/*.h)
	gcc -O2 -ffp-contract=off -Wall -Wno-missing-braces -Wno-unused-function -Wno-misleading-indentation -shared -fPIC -o $@ $(OSRC) -lm

# CPU-baseline build of the oracle for the machine it runs on (bench.py builds it ON the box: -march=native code must not
# travel between hosts); same sources, same results, -O3.
ORACLE_NATIVE := oracle/_native/liboracle_native.so
oracle-native: $(ORACLE_NATIVE)
$(ORACLE_NATIVE): $(OSRC) $(HDR) $(wildcard oracle/*.h)
	mkdir -p oracle/_native
	gcc -O3 -march=native -ffp-contract=off -Wall -Wno-missing-braces -Wno-unused-function -Wno-misleading-indentation -Wno-maybe-uninitialized -shared -fPIC -o $@ $(OSRC) -lm

clean:
	rm -f $(LIB) $(ORACLE)
.PHONY: all lib oracle oracle-native clean
