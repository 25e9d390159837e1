#!/bin/bash
for v in 16 32 33 34 35 36 37 64 66 1; do timeout 20 scripts/probe/umma_mn_probe $v 2>&1 | tail -1; done
