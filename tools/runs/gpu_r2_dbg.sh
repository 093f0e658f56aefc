#!/bin/bash
timeout 1800 python -m pytest tests -q -m gpu -x 2>&1 | grep -E "^FAILED|^E  |Error|assert" | head -20
