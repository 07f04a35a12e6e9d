// TEST INFRASTRUCTURE: included by share_mem.h, unused (the scoped_lock lines are commented out).
#pragma once
