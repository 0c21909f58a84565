cd $GRAFT_REPO_ROOT
echo "== S=8 full"; WSR_SPLITK_FORCE=256,8 python tools/prof_conv.py 8 2>&1 | head -3
for d in 8 16 32 24 48 56 1 2; do echo "== S=8 dbg=$d"; WSR_TC_DBG=$d WSR_SPLITK_FORCE=256,8 python tools/prof_conv.py 8 2>&1 | head -3; done
echo "== nosplit bn=256 dbg=1 (no epilogue)"; WSR_TC_DBG=1 WSR_SPLITK_FORCE=256,1 python tools/prof_conv.py 8 2>&1 | head -3
echo "== nosplit bn=256 dbg=2 (no MMA)"; WSR_TC_DBG=2 WSR_SPLITK_FORCE=256,1 python tools/prof_conv.py 8 2>&1 | head -3
echo "== nosplit bn=64 dbg=2 (no MMA)"; WSR_TC_DBG=2 WSR_NO_SPLITK=1 python tools/prof_conv.py 8 2>&1 | head -3
