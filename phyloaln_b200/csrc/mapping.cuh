// mapping.cuh -- hit -> reference-column mapping kernel (placeholder).
#pragma once
