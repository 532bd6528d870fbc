cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests -x -q -m gpu -k multipass 2>&1 | grep -E "^E " | cut -c1-400 | tail -15
