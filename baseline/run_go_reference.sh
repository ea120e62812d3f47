#!/usr/bin/env bash
# CPU baseline with the REAL reference (option 1): go-pocket-tts' own stage
# profiler on the native-safetensors backend, AVX2/FMA dot on (dot_amd64.go:9-19), all host cores.
# Needs a Go toolchain (>= go.mod's version), the reference checkout and its model files; none of
# these exist in the build image or on the GPU boxes, so bench.py's reference arm times the C
# restatement (oracle/) instead and says so.  Run this wherever Go is available and put the printed
# `rtf` / `avg_generate_ms` beside bench.py's cpu_baseline.
#
#   baseline/run_go_reference.sh [/path/to/go-pocket-tts] ["text to synthesise"]
set -euo pipefail
REF="${1:-/root/reference}"
TEXT="${2:-Hello world.}"
if ! command -v go >/dev/null 2>&1; then
  echo "run_go_reference: no Go toolchain on PATH (go: command not found); the reference cannot be built here." >&2
  echo "run_go_reference: bench.py --impl reference reports the C restatement of the same path instead." >&2
  exit 3
fi
if [ ! -f "$REF/go.mod" ]; then
  echo "run_go_reference: $REF is not a go-pocket-tts checkout" >&2
  exit 2
fi
CORES="$(nproc)"
cd "$REF"
# single-stream real-time factor (internal/bench/stageprof/stageprof.go:137-162 prints audio_ms,
# avg_generate_ms, rtf = avg_total_ms / audio_ms)
go run ./bench/stageprof --text "$TEXT" --runs 5 --warmup 1 \
  --runtime-workers "$CORES" --conv-workers "$CORES"
echo "cores: $CORES"
echo "note: throughput configs (C3-C5) = $CORES concurrent GenerateAudio goroutines; the AR steps of"
echo "      concurrent requests are serialised by rngMu (runtime_native_safetensors.go:161-170)."
