/* stand-in for VTK's vtkStructuredGrid.h: see vtk_standin.h (test infrastructure) */
#include "vtk_standin.h"
