# Same targets as the reference makefile (all q1 q2 q3 benchmark run-q1 .. run-all analyze clean
# clean-all).  The drivers are host-only C++ linked against the engine's C-ABI library; the CUDA
# code is built once into libhw1f.so (sm_100a).
CXX      ?= g++
PKG      := monte-carlo-simulation-of-hull-white-model-and-sensitivities-computation_b200
LIBDIR   := $(PKG)/lib
CXXFLAGS := -O2 -std=c++17 -Wall -I include
LDFLAGS  := -L$(LIBDIR) -lhw1f -Wl,-rpath,'$$ORIGIN/../$(LIBDIR)'
SRC      := src
BIN      := bin
# the reference pins CUDA_VISIBLE_DEVICES=2 (makefile:27-39), which hides every GPU on a box with
# fewer than three; override with `make run-q1 GPU=2` to reproduce that
GPU      ?= 0

all: q1 q2 q3

$(LIBDIR)/libhw1f.so: $(PKG)/csrc/*.cu $(PKG)/csrc/*.cuh $(PKG)/csrc/*.cpp $(PKG)/csrc/*.hpp include/hw1f.h
	python3 $(PKG)/build.py --force

q1: $(SRC)/1_bond_pricing.cpp $(LIBDIR)/libhw1f.so
	@mkdir -p $(BIN) data plots
	$(CXX) $(CXXFLAGS) $< -o $(BIN)/$@ $(LDFLAGS)

q2: $(SRC)/2_option_pricing.cpp $(LIBDIR)/libhw1f.so
	@mkdir -p $(BIN) data plots
	$(CXX) $(CXXFLAGS) $< -o $(BIN)/$@ $(LDFLAGS)

q3: $(SRC)/3_sensitivity_analysis.cpp $(LIBDIR)/libhw1f.so
	@mkdir -p $(BIN) data plots
	$(CXX) $(CXXFLAGS) $< -o $(BIN)/$@ $(LDFLAGS)

benchmark: $(SRC)/benchmark_reductions.cpp $(LIBDIR)/libhw1f.so
	@mkdir -p $(BIN) data plots
	$(CXX) $(CXXFLAGS) $< -o $(BIN)/$@ $(LDFLAGS)

run-q1: q1
	@mkdir -p data plots
	CUDA_VISIBLE_DEVICES=$(GPU) ./$(BIN)/q1

run-q2: q2
	@mkdir -p data plots
	CUDA_VISIBLE_DEVICES=$(GPU) ./$(BIN)/q2

run-q3: q3
	@mkdir -p data plots
	CUDA_VISIBLE_DEVICES=$(GPU) ./$(BIN)/q3

run-benchmark: benchmark
	@mkdir -p data plots
	CUDA_VISIBLE_DEVICES=$(GPU) ./$(BIN)/benchmark

run-all: run-q1 run-q2 run-q3

# analyze.py belongs to the reference (plots from data/*); set ANALYZE to its path to run it
ANALYZE ?= analyze.py
analyze: run-all run-benchmark
	@if [ -f $(ANALYZE) ]; then python3 $(ANALYZE); else echo "analyze.py not present: data/ holds the files it reads"; fi

clean:
	rm -rf $(BIN) data/*.json data/*.csv data/summary.txt plots/*.png

clean-all:
	rm -rf $(BIN) data plots

.PHONY: all clean clean-all run-q1 run-q2 run-q3 run-benchmark run-all analyze
