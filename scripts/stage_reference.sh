#!/bin/bash
# Stage a copy of the reference's app/ for ONE gpurun call of tests/test_reference_handlers_gpu.py
# (the GPU box has no /root/reference).  The copy lives under the git-ignored oracle/_ref/ and is
# removed again by `scripts/stage_reference.sh clean` right after the call: reference sources never
# enter the repository.
set -e
cd "$(dirname "$0")/.."
if [ "$1" = "clean" ]; then rm -rf oracle/_ref/reference; echo "removed oracle/_ref/reference"; exit 0; fi
mkdir -p oracle/_ref/reference/app
cp /root/reference/app/main.py /root/reference/app/embedding_gen.py oracle/_ref/reference/app/
echo "staged $(ls oracle/_ref/reference/app)"
