# Host-only drivers over the engine's C ABI (libhw1f.so, built once for sm_100a by build.py).
# Target names follow the reference makefile so that existing scripts keep working:
#   all | q1 q2 q3 benchmark | run-q1 run-q2 run-q3 run-benchmark run-all | analyze | clean clean-all
PKG      := monte-carlo-simulation-of-hull-white-model-and-sensitivities-computation_b200
ENGINE   := $(PKG)/lib/libhw1f.so
CXX      ?= g++
CXXFLAGS := -O2 -std=c++17 -Wall -I include
LDLIBS   := -L$(PKG)/lib -lhw1f -Wl,-rpath,'$$ORIGIN/../$(PKG)/lib'
OUTDIRS  := bin data plots

# driver name -> source stem
stem_q1        := 1_bond_pricing
stem_q2        := 2_option_pricing
stem_q3        := 3_sensitivity_analysis
stem_benchmark := benchmark_reductions
DRIVERS        := q1 q2 q3 benchmark

# the reference pins CUDA_VISIBLE_DEVICES=2, which hides every GPU on a box with fewer than three;
# `make run-q1 GPU=2` reproduces that
GPU ?= 0

all: q1 q2 q3

$(ENGINE): $(wildcard $(PKG)/csrc/*) include/hw1f.h
	python3 $(PKG)/build.py --force

$(OUTDIRS):
	@mkdir -p $@

.SECONDEXPANSION:
$(DRIVERS): %: src/$$(stem_$$*).cpp include/hw1f.h include/hw1f_driver.hpp $(ENGINE) | $(OUTDIRS)
	$(CXX) $(CXXFLAGS) $< -o bin/$@ $(LDLIBS)

# host overhead of the C ABI measured from C++ (tools/api_overhead.cpp)
api_overhead: tools/api_overhead.cpp include/hw1f.h include/hw1f_driver.hpp $(ENGINE) | $(OUTDIRS)
	$(CXX) $(CXXFLAGS) $< -o bin/$@ $(LDLIBS)

$(addprefix run-,$(DRIVERS)): run-%: % | $(OUTDIRS)
	CUDA_VISIBLE_DEVICES=$(GPU) ./bin/$*

run-all: run-q1 run-q2 run-q3

# analyze.py (matplotlib plots from data/*) belongs to the reference; point ANALYZE at it to run it
ANALYZE ?= analyze.py
analyze: run-all run-benchmark
	@if [ -f $(ANALYZE) ]; then python3 $(ANALYZE); else echo "analyze.py not present: data/ holds the files it reads"; fi

clean:
	$(RM) -r bin $(addprefix data/,*.json *.csv summary.txt) plots/*.png

clean-all:
	$(RM) -r $(OUTDIRS)

.PHONY: all run-all analyze clean clean-all api_overhead $(addprefix run-,$(DRIVERS))
