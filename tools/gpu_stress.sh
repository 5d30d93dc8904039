#!/bin/bash
timeout 600 python tools/stress_nms_cluster.py 2>&1 | tail -5; echo "exit $?"
