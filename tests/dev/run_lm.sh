for s in 3 4 5; do timeout 600 python tests/dev/fuzz_logmel.py $s 2>&1 | tail -1; done
for s in 1 2; do timeout 600 python tests/dev/fuzz_pcm.py $s 2>&1 | tail -1; done
