#!/bin/bash
timeout 600 python tools/preprocess_order_probe.py 2>&1 | tail -20
