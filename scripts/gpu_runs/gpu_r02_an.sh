cd $GRAFT_REPO_ROOT
timeout 300 python scripts/e2e_timeline.py --out gpurun_out/r02_tl3.json --window 2.4,4.4 > /dev/null 2> gpurun_out/r02_tl3.err
rm -f gpurun_out/*_chrome.json
grep -v -i warn gpurun_out/r02_tl3.err | head -150
