#!/bin/bash
# Round 2, GPU call 17 = call 15 (validation of the device pair support, flat average pool, allele tile kernel; the product CLI's
# wall clock) followed by call 16 (halo rule 2 / CTA-pair rule 3 experiments), in one box session.
bash tools/gpu_r2_call15.sh
bash tools/gpu_r2_call16.sh
