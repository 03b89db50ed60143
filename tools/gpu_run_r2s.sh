O=gpurun_out/r2s; mkdir -p $O
for args in "44100 48000 2 65536 float" "44100 48000 2 65536 double" "44100 48000 2 16384 float" "44100 48000 2 262144 float" "192000 44100 8 65536 double" "384000 48000 8 65536 float"; do
  B200RATE_TRACE_STREAM=1 tools/bin/stream_lat $args >> $O/stream_lat.txt 2>&1
done
cat $O/stream_lat.txt
