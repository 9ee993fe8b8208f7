for mode in ordered free; do
  for mib in 384 768; do
    if [ $mode = free ]; then export ATDE_HOST_FREE=1; else unset ATDE_HOST_FREE; fi
    echo "mode=$mode"; python profiles/tools/exp_host_path.py $mib 2>&1 | tail -1 | cut -c1-120
  done
done
