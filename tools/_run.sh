timeout 600 python -m pytest tests -m gpu -x -q -k "stream_form" 2>&1 | tail -2
timeout 600 python tools/tune.py --tag tail --sweep "TAIL=0;TAIL=15;TAIL=30;TAIL=50;TAIL=15,L=128;TAIL=30,L=128;TAIL=50,L=256" 2>&1 | tee gpurun_out/st_tune17.log
