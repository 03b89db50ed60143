for g in 4 5 6; do echo GROUPS=$g; B200RATE_DFT_GROUPS=$g python tools/stage_probe.py 2>&1 | grep -v "stage " | grep -E "^48000->44100 2ch x256|^384000|^44100->96000" | cut -c1-70; done
