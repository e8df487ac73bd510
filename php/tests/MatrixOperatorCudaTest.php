<?php
namespace RindowTest\Math\Matrix\MatrixOperatorCudaTest;

use Rindow\Math\Matrix\MatrixOperatorCuda;
use Rindow\Math\Matrix\Drivers\Service;
use Rindow\Math\Matrix\Drivers\MatlibCuda\MatlibCuda;
// the reference's test classes are not autoloaded (cf. LinearAlgebraPHPModeTest.php:12-14)
if(!class_exists('RindowTest\Math\Matrix\MatrixOperatorTest\MatrixOperatorTest')) {
    require_once __DIR__.'/MatrixOperatorTest.php';
}
use RindowTest\Math\Matrix\MatrixOperatorTest\MatrixOperatorTest as ORGTest;

/**
 * MatrixOperatorTest over the B200 driver (cf. MatrixOperatorPhpModeTest.php:14-25).  The operator is
 * MatrixOperatorCuda, so the reference's testSumWithAxis / testMaxWithAxis / testArgMaxWithAxis / testMinWithAxis /
 * testArgMinWithAxis expectations (:544-560, :620-665, :687-730, :838-913) are met by the reduce kernels instead of one
 * driver call per output element.
 */
class MatrixOperatorCudaTest extends ORGTest
{
    public function newMatrixOperator()
    {
        $service = new MatlibCuda();
        if($service->serviceLevel()<Service::LV_ADVANCED) {
            $this->markTestSkipped('librindow_b200.so / B200 not available');
        }
        return new MatrixOperatorCuda(service:$service);
    }
}
