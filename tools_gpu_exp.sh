#!/bin/bash
for B in 64 256; do B=$B timeout 300 python tools/update_one.py; done
