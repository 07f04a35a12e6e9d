// TEST INFRASTRUCTURE: share_mem.h includes this header but uses nothing from it.
#pragma once
