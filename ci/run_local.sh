#!/bin/bash
# Local equivalent of the CI job: build + every test that does not need a GPU.
set -e
cd "$(dirname "$0")/.."
make -j"$(nproc)"
make test
python -m pytest tests -q -m "not gpu"
