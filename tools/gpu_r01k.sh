#!/bin/bash
timeout -s KILL 300 python tools/qtc_probe.py int4
timeout -s KILL 300 python tools/qtc_probe.py sq8
