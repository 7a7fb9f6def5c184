set -x
T=r3o
timeout -s KILL 300 compute-sanitizer --tool racecheck --racecheck-report all python __graft_entry__.py smoke > gpurun_out/${T}_racecheck_full.txt 2>&1
grep -c "hazard" gpurun_out/${T}_racecheck_full.txt
grep -E "Error:|Warning:|hazards\]|Write access|Read access" gpurun_out/${T}_racecheck_full.txt | sort | uniq -c | sort -rn | head -40
