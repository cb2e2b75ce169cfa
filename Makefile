# Build libthunder_b200.so (sm_100a only) and the C oracle.  `python -c "import __graft_entry__ as g; g.build()"`
# runs the same commands.
NVCC      ?= nvcc
CXX       := g++
ARCH      := -gencode arch=compute_100a,code=sm_100a
NVFLAGS   := $(ARCH) -O3 -lineinfo -std=c++17 -Xcompiler -fPIC,-fopenmp -Xptxas -v
CSRC      := thunder_b200/csrc
OUT       := thunder_b200/lib/libthunder_b200.so
OBJS      := build/thb_api.o build/thb_pf.o build/thb_reco.o build/thb_comm.o
HDRS      := $(wildcard $(CSRC)/*.cuh $(CSRC)/*.h include/*.h)

IFACE     := thunder_b200/lib/libthb_interface.so

all: $(OUT) $(IFACE) oracle

build/%.o: $(CSRC)/%.cu $(HDRS)
	@mkdir -p build
	$(NVCC) $(NVFLAGS) -Iinclude -c $< -o $@ 2> build/$*.ptxas.log || (cat build/$*.ptxas.log; false)

# the particle filter is compiled WITHOUT fused multiply-add contraction: the reference's ACG inference (fixed-point iteration on
# a 4x4 matrix of condition 1e4 .. 1e9, cofactor inverse) amplifies last-bit differences, and the parity tests replay the
# reference's Particle class draw by draw - its build (gcc -O2 -mavx) does not contract either
build/thb_pf.o: $(CSRC)/thb_pf.cu $(HDRS)
	@mkdir -p build
	$(NVCC) $(NVFLAGS) -fmad=false -Iinclude -c $< -o $@ 2> build/thb_pf.ptxas.log || (cat build/thb_pf.ptxas.log; false)

build/thb_comm.o: $(CSRC)/thb_comm.cpp $(HDRS)
	@mkdir -p build
	$(NVCC) $(NVFLAGS) -Iinclude -x cu -c $< -o $@ 2> build/thb_comm.ptxas.log || (cat build/thb_comm.ptxas.log; false)

$(OUT): $(OBJS)
	@mkdir -p thunder_b200/lib
	$(NVCC) $(ARCH) -shared -Xcompiler -fPIC -o $@ $(OBJS) -Xlinker --no-as-needed -lgomp -ldl -lcufft -Xlinker -rpath -Xlinker /usr/local/cuda/lib64

# host-side mirror of the reference's accelerator seam (C++), on top of the C ABI
$(IFACE): thunder_b200/host/Interface.cpp thunder_b200/host/Interface.h include/thunder_b200.h $(OUT)
	$(CXX) -O2 -std=c++17 -fPIC -shared -Iinclude -o $@ thunder_b200/host/Interface.cpp -Lthunder_b200/lib -lthunder_b200 -Wl,-rpath,'$$ORIGIN'

oracle: oracle/_port/libthb_oracle.so

oracle/_port/libthb_oracle.so: oracle/thb_oracle.c
	@mkdir -p oracle/_port
	gcc -O2 -fPIC -shared -std=gnu99 -o $@ $< -lm

clean:
	rm -rf build $(OUT) oracle/_port

.PHONY: all oracle clean
