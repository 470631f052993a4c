#!/bin/bash
# What is run on the B200 box to validate a build (round 2): full GPU suite, smoke, bench (both depth stages), reference arm.
exec bash "$(dirname "$0")/r2_call41.sh"
